/* fakegl.c -- TEST INFRASTRUCTURE (oracle).
 *
 * A just-big-enough OpenGL state machine for the reference's horizonator-lib.c to run
 * unmodified (built by oracle/Makefile into oracle/_ref/libhorizonator_ref.so).  It keeps
 * buffer objects, one vertex attribute, one program's uniforms by name, renderbuffers and
 * framebuffer objects; glDrawElements() runs the GL-pipeline restatement of
 * oracle/gl_pipeline.c (the three shaders of the reference restated in C plus the
 * fixed-function rules F1-F9).  With this, every line of the reference's host code -- DEM
 * stitching, vertex/index buffer fill, uniform math, auto viewer height, read-back, vertical
 * flip, depth->range -- executes as written; only the driver is a restatement.
 *
 * The draw honours exactly the state horizonator-lib.c sets (depth test + back-face culling
 * on, clear colour 0,0,1, full-buffer viewport) and aborts if it finds anything else.
 */
#include <epoxy/gl.h>
#include <GL/freeglut.h>
#include <FreeImage.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../gl_pipeline.h"

#define DIE(...) do { fprintf(stderr, "fakegl: " __VA_ARGS__); fprintf(stderr, "\n"); abort(); } while(0)

/* ---- objects ---- */
#define MAXOBJ 64
typedef struct { void* data; size_t size; } buffer_t;
typedef struct { GLenum format; int w, h; } renderbuffer_t;
typedef struct { GLuint color_rb, depth_rb; glp_framebuffer_t fb; int allocated; } framebuffer_t;
typedef struct { char name[48]; float f; int i; } uniform_t;

static buffer_t       g_buffers[MAXOBJ];       static GLuint g_nbuffers = 0;
static renderbuffer_t g_rbs[MAXOBJ];           static GLuint g_nrbs     = 0;
static framebuffer_t  g_fbos[MAXOBJ];          static GLuint g_nfbos    = 0;  /* [0] = window */
static uniform_t      g_uniforms[MAXOBJ];      static int    g_nuniforms = 0;
static GLuint g_nshaders = 0, g_nprograms = 0, g_nvaos = 0, g_ntextures = 0;

static GLuint g_bound_array = 0, g_bound_element = 0, g_bound_rb = 0, g_bound_fbo = 0;
static GLuint g_attr0_buffer = 0;
static int    g_attr0_ok = 0;
static int    g_depth_test = 0, g_cull_face = 0;
static float  g_clear[4] = {0, 0, 0, 0};
static int    g_viewport[4] = {0, 0, 1024, 1024};
static int    g_pack_alignment = 4;
static int    g_window_w = 1024, g_window_h = 1024;
static int    g_threads = 1;

void fakegl_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

/* Test helper, not GL: the whole n x n DEM square through the reference's OWN sampler (dem.c:264-309, compiled
 * unmodified into this library), one call per cell in the order horizonator-lib.c:435-439 makes them.  out[j*n+i]. */
#include <stdint.h>
struct horizonator_dem_context_t;
int16_t horizonator_dem_sample(const struct horizonator_dem_context_t* ctx, int i, int j);
void ref_sample_square(const void* dem_ctx, int n, int16_t* out)
{
    #pragma omp parallel for schedule(static) num_threads(g_threads)
    for(int j = 0; j < n; j++)
        for(int i = 0; i < n; i++)
            out[(size_t)j * n + i] = horizonator_dem_sample((const struct horizonator_dem_context_t*)dem_ctx, i, j);
}

/* ---- misc ---- */
GLenum glGetError(void) { return GL_NO_ERROR; }
const unsigned char* glGetString(GLenum name)
{
    if(name == GL_VERSION) return (const unsigned char*)"4.2 (fakegl: CPU restatement, oracle only)";
    return (const unsigned char*)"";
}
void glEnable(GLenum cap)
{
    if(cap == GL_DEPTH_TEST) g_depth_test = 1;
    else if(cap == GL_CULL_FACE) g_cull_face = 1;
    else DIE("glEnable(%#x) not modelled", cap);
}
void glClearColor(GLfloat r, GLfloat g, GLfloat b, GLfloat a) { g_clear[0]=r; g_clear[1]=g; g_clear[2]=b; g_clear[3]=a; }
void glViewport(GLint x, GLint y, GLsizei w, GLsizei h) { g_viewport[0]=x; g_viewport[1]=y; g_viewport[2]=w; g_viewport[3]=h; }
void glGetIntegerv(GLenum pname, GLint* data)
{
    if(pname == GL_VIEWPORT) memcpy(data, g_viewport, sizeof(g_viewport));
    else DIE("glGetIntegerv(%#x) not modelled", pname);
}
void glPixelStorei(GLenum pname, GLint param) { if(pname == GL_PACK_ALIGNMENT) g_pack_alignment = param; }
void glDrawBuffer(GLenum buf) { (void)buf; }

/* ---- buffers / vertex arrays ---- */
void glGenVertexArrays(GLsizei n, GLuint* ids) { for(int k=0;k<n;k++) ids[k] = ++g_nvaos; }
void glBindVertexArray(GLuint id) { (void)id; }
void glGenBuffers(GLsizei n, GLuint* ids)
{
    for(int k=0;k<n;k++) { if(g_nbuffers+1 >= MAXOBJ) DIE("too many buffers"); ids[k] = ++g_nbuffers; }
}
static GLuint* bound_buffer(GLenum target)
{
    if(target == GL_ARRAY_BUFFER) return &g_bound_array;
    if(target == GL_ELEMENT_ARRAY_BUFFER) return &g_bound_element;
    DIE("buffer target %#x not modelled", target);
}
void glBindBuffer(GLenum target, GLuint id) { *bound_buffer(target) = id; }
void glBufferData(GLenum target, GLsizeiptr size, const void* data, GLenum usage)
{
    (void)usage;
    buffer_t* b = &g_buffers[*bound_buffer(target)];
    free(b->data);
    b->data = malloc(size); b->size = size;
    if(!b->data) DIE("out of memory for a %zd-byte buffer", (ptrdiff_t)size);
    if(data) memcpy(b->data, data, size);
}
void* glMapBuffer(GLenum target, GLenum access) { (void)access; return g_buffers[*bound_buffer(target)].data; }
GLboolean glUnmapBuffer(GLenum target) { (void)target; return GL_TRUE; }
void glEnableVertexAttribArray(GLuint index) { if(index != 0) DIE("attribute %u not modelled", index); }
void glVertexAttribPointer(GLuint index, GLint size, GLenum type, GLboolean normalized,
                           GLsizei stride, const void* pointer)
{
    /* horizonator-lib.c:424: attribute 0 = 3 x GL_SHORT, not normalised, tightly packed */
    g_attr0_ok = (index == 0 && size == 3 && type == GL_SHORT && !normalized && stride == 0 && pointer == NULL);
    g_attr0_buffer = g_bound_array;
}

/* ---- program / uniforms ---- */
GLuint glCreateProgram(void) { return ++g_nprograms; }
GLuint glCreateShader(GLenum type) { (void)type; return ++g_nshaders; }
void glShaderSource(GLuint s, GLsizei c, const GLchar** str, const GLint* len) { (void)s;(void)c;(void)str;(void)len; }
void glCompileShader(GLuint s) { (void)s; }
void glGetShaderInfoLog(GLuint s, GLsizei n, GLsizei* len, GLchar* log) { (void)s; if(n>0) log[0]=0; if(len) *len=0; }
void glAttachShader(GLuint p, GLuint s) { (void)p;(void)s; }
void glLinkProgram(GLuint p) { (void)p; }
void glUseProgram(GLuint p) { (void)p; }
void glGetProgramInfoLog(GLuint p, GLsizei n, GLsizei* len, GLchar* log) { (void)p; if(n>0) log[0]=0; if(len) *len=0; }
GLint glGetUniformLocation(GLuint program, const GLchar* name)
{
    (void)program;
    for(int k=0;k<g_nuniforms;k++) if(!strcmp(g_uniforms[k].name, name)) return k;
    if(g_nuniforms >= MAXOBJ) DIE("too many uniforms");
    snprintf(g_uniforms[g_nuniforms].name, sizeof(g_uniforms[0].name), "%s", name);
    return g_nuniforms++;
}
void glUniform1f(GLint loc, GLfloat v) { if(loc >= 0 && loc < g_nuniforms) g_uniforms[loc].f = v; }
void glUniform1i(GLint loc, GLint v)   { if(loc >= 0 && loc < g_nuniforms) g_uniforms[loc].i = v; }
void glGetUniformfv(GLuint program, GLint loc, GLfloat* params)
{
    (void)program;
    if(loc < 0 || loc >= g_nuniforms) DIE("glGetUniformfv: bad location %d", loc);
    *params = g_uniforms[loc].f;
}
static float uniform_f(const char* name)
{
    for(int k=0;k<g_nuniforms;k++) if(!strcmp(g_uniforms[k].name, name)) return g_uniforms[k].f;
    DIE("uniform '%s' was never created", name);
}
static int uniform_i(const char* name)
{
    for(int k=0;k<g_nuniforms;k++) if(!strcmp(g_uniforms[k].name, name)) return g_uniforms[k].i;
    DIE("uniform '%s' was never created", name);
}

/* ---- framebuffers ---- */
void glGenFramebuffers(GLsizei n, GLuint* ids)
{
    for(int k=0;k<n;k++) { if(g_nfbos+1 >= MAXOBJ) DIE("too many FBOs"); ids[k] = ++g_nfbos; }
}
void glBindFramebuffer(GLenum target, GLuint id) { (void)target; g_bound_fbo = id; }
void glGenRenderbuffers(GLsizei n, GLuint* ids)
{
    for(int k=0;k<n;k++) { if(g_nrbs+1 >= MAXOBJ) DIE("too many renderbuffers"); ids[k] = ++g_nrbs; }
}
void glBindRenderbuffer(GLenum target, GLuint id) { (void)target; g_bound_rb = id; }
void glRenderbufferStorage(GLenum target, GLenum fmt, GLsizei w, GLsizei h)
{
    (void)target;
    if(fmt != GL_RGB && fmt != GL_DEPTH_COMPONENT) DIE("renderbuffer format %#x not modelled", fmt);
    g_rbs[g_bound_rb].format = fmt; g_rbs[g_bound_rb].w = w; g_rbs[g_bound_rb].h = h;
}
void glFramebufferRenderbuffer(GLenum target, GLenum attachment, GLenum rbtarget, GLuint rb)
{
    (void)target; (void)rbtarget;
    if(attachment == GL_COLOR_ATTACHMENT0)    g_fbos[g_bound_fbo].color_rb = rb;
    else if(attachment == GL_DEPTH_ATTACHMENT) g_fbos[g_bound_fbo].depth_rb = rb;
    else DIE("attachment %#x not modelled", attachment);
}
GLenum glCheckFramebufferStatus(GLenum target) { (void)target; return GL_FRAMEBUFFER_COMPLETE; }

static glp_framebuffer_t* current_fb(void)
{
    framebuffer_t* f = &g_fbos[g_bound_fbo];
    if(!f->allocated)
    {
        int w, h;
        if(g_bound_fbo == 0) { w = g_window_w; h = g_window_h; }
        else
        {
            const renderbuffer_t* c = &g_rbs[f->color_rb], *d = &g_rbs[f->depth_rb];
            if(!f->color_rb || !f->depth_rb || c->format != GL_RGB || d->format != GL_DEPTH_COMPONENT ||
               c->w != d->w || c->h != d->h)
                DIE("FBO %u is not RGB colour + DEPTH_COMPONENT depth of equal size", g_bound_fbo);
            w = c->w; h = c->h;
        }
        if(!glp_framebuffer_alloc(&f->fb, w, h)) DIE("out of memory for a %dx%d framebuffer", w, h);
        glp_clear(&f->fb);
        f->allocated = 1;
    }
    return &f->fb;
}

void glClear(GLbitfield mask)
{
    if(mask != (GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT)) DIE("glClear(%#x) not modelled", mask);
    if(!(g_clear[0] == 0.f && g_clear[1] == 0.f && g_clear[2] == 1.f))
        DIE("clear colour is not (0,0,1): gl_pipeline.c restates horizonator-lib.c:185 only");
    glp_clear(current_fb());
}

void glDrawElements(GLenum mode, GLsizei count, GLenum type, const void* indices)
{
    glp_framebuffer_t* fb = current_fb();
    if(mode != GL_TRIANGLES || type != GL_UNSIGNED_INT || indices != NULL) DIE("draw call not modelled");
    if(!g_depth_test || !g_cull_face) DIE("depth test and face culling must both be enabled");
    if(!g_attr0_ok) DIE("vertex attribute 0 is not 3 x GL_SHORT");
    if(g_viewport[0] || g_viewport[1] || g_viewport[2] != fb->width || g_viewport[3] != fb->height)
        DIE("viewport must cover the whole %dx%d target", fb->width, fb->height);
    if(uniform_i("NtilesX") != 0) DIE("textured rendering is out of scope");

    glp_uniforms_t u;
    u.viewer_cell_i  = uniform_f("viewer_cell_i");
    u.viewer_cell_j  = uniform_f("viewer_cell_j");
    u.viewer_z       = uniform_f("viewer_z");
    u.DEG_PER_CELL   = uniform_f("DEG_PER_CELL");
    u.cos_viewer_lat = uniform_f("cos_viewer_lat");
    u.az_deg0        = uniform_f("az_deg0");
    u.az_deg1        = uniform_f("az_deg1");
    u.aspect         = uniform_f("aspect");
    u.znear          = uniform_f("znear");
    u.zfar           = uniform_f("zfar");
    u.curvature      = 0.0f;     /* the reference has no such uniform */
    u.seam_wrap      = 0;        /* nor this option */
    u.znear_color    = uniform_f("znear_color");
    u.zfar_color     = uniform_f("zfar_color");

    const buffer_t* vb = &g_buffers[g_attr0_buffer], *ib = &g_buffers[g_bound_element];
    if((size_t)count * sizeof(GLuint) > ib->size) DIE("index buffer too small");
    glp_draw_triangles(fb, &u, (const int16_t*)vb->data, (int64_t)(vb->size / (3 * sizeof(GLshort))),
                       (const uint32_t*)ib->data, (int64_t)count / 3, 0, g_threads);
}

void glReadPixels(GLint x, GLint y, GLsizei w, GLsizei h, GLenum format, GLenum type, void* pixels)
{
    glp_framebuffer_t* fb = current_fb();
    if(x < 0 || y < 0 || x + w > fb->width || y + h > fb->height) DIE("glReadPixels out of bounds");
    if(format == GL_BGR && type == GL_UNSIGNED_BYTE)
    {
        if(x || y || w != fb->width || h != fb->height) DIE("partial colour read-back not modelled");
        if(g_pack_alignment != 1 && (w * 3) % g_pack_alignment) DIE("row padding not modelled");
        glp_read_bgr(fb, (uint8_t*)pixels);
    }
    else if(format == GL_DEPTH_COMPONENT && type == GL_FLOAT)
        glp_read_depth_float(fb, x, y, w, h, (float*)pixels);
    else DIE("glReadPixels(%#x,%#x) not modelled", format, type);
}

/* ---- textures: out of scope, never reached untextured ---- */
void glGenTextures(GLsizei n, GLuint* ids) { for(int k=0;k<n;k++) ids[k] = ++g_ntextures; }
void glActiveTextureARB(GLenum t) { (void)t; }
void glBindTexture(GLenum t, GLuint id) { (void)t;(void)id; }
void glTexParameteri(GLenum t, GLenum p, GLint v) { (void)t;(void)p;(void)v; }
void glTexImage2D(GLenum t, GLint l, GLint f, GLsizei w, GLsizei h, GLint b, GLenum fo, GLenum ty, const void* p)
{ (void)t;(void)l;(void)f;(void)w;(void)h;(void)b;(void)fo;(void)ty;(void)p; DIE("textures are out of scope"); }
void glTexSubImage2D(GLenum t, GLint l, GLint xo, GLint yo, GLsizei w, GLsizei h, GLenum fo, GLenum ty, const void* p)
{ (void)t;(void)l;(void)xo;(void)yo;(void)w;(void)h;(void)fo;(void)ty;(void)p; DIE("textures are out of scope"); }

/* ---- GLUT: no window system ---- */
void glutInitContextFlags(int f) { (void)f; }
void glutInitContextVersion(int a, int b) { (void)a;(void)b; }
void glutInitContextProfile(int p) { (void)p; }
void glutInit(int* argc, char** argv) { (void)argc;(void)argv; }
void glutInitDisplayMode(unsigned int m) { (void)m; }
void glutInitWindowSize(int w, int h) { g_window_w = w; g_window_h = h; }
int  glutCreateWindow(const char* t) { (void)t; static int n = 0; return ++n; }
void glutHideWindow(void) {}
int  glutExtensionSupported(const char* e) { (void)e; return 1; }
void glutSetWindow(int id) { (void)id; }
void glutDestroyWindow(int id) { (void)id; }
void glutExit(void) {}

/* ---- FreeImage: texture path only ---- */
FREE_IMAGE_FORMAT FreeImage_GetFileType(const char* f, int s) { (void)f;(void)s; return FIF_UNKNOWN; }
FIBITMAP* FreeImage_Load(FREE_IMAGE_FORMAT a, const char* f, int fl) { (void)a;(void)f;(void)fl; return NULL; }
FREE_IMAGE_COLOR_TYPE FreeImage_GetColorType(FIBITMAP* d) { (void)d; return 0; }
FIBITMAP* FreeImage_ConvertTo24Bits(FIBITMAP* d) { (void)d; return NULL; }
void FreeImage_Unload(FIBITMAP* d) { (void)d; }
unsigned FreeImage_GetWidth(FIBITMAP* d) { (void)d; return 0; }
unsigned FreeImage_GetHeight(FIBITMAP* d) { (void)d; return 0; }
unsigned FreeImage_GetBPP(FIBITMAP* d) { (void)d; return 0; }
unsigned FreeImage_GetPitch(FIBITMAP* d) { (void)d; return 0; }
BYTE* FreeImage_GetBits(FIBITMAP* d) { (void)d; return NULL; }
