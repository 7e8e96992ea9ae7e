/* glut_glx.c -- TEST INFRASTRUCTURE (oracle).
 *
 * The handful of freeglut calls the reference's horizonator-lib.c makes (horizonator-lib.c:123-176, 665-687),
 * implemented on GLX pbuffers so that the reference, compiled unmodified, runs on a REAL OpenGL driver: the
 * Mesa llvmpipe libGL that ships inside the image with Nsight Compute, talking to the display-less Xlib of
 * fakex11.c.  glutCreateWindow() makes a core-profile context of the version the reference asked for
 * (glutInitContextVersion(4,2)) current on a pbuffer of the requested window size; everything else is
 * bookkeeping.  The gl*() calls of the reference bind straight to Mesa's libGL: this file has no GL logic.
 *
 * llvmpipe of that vintage (Mesa 18.1) is an OpenGL 3.3 implementation; the reference's shaders say
 * "#version 420" but use nothing newer than GLSL 3.30, so the version is raised with Mesa's own override
 * variables (MESA_GL_VERSION_OVERRIDE / MESA_GLSL_VERSION_OVERRIDE) -- the reference's sources stay untouched.
 */
#include <GL/freeglut.h>

#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef struct FakeDisplay Display;
typedef unsigned long XID;
typedef void* GLXFBConfig;
typedef void* GLXContext;
typedef void (*glx_proc_t)(void);

extern Display* XOpenDisplay(const char*);
extern int XFree(void*);
extern GLXFBConfig* glXChooseFBConfig(Display*, int screen, const int* attribs, int* n);
extern XID  glXCreatePbuffer(Display*, GLXFBConfig, const int* attribs);
extern void glXDestroyPbuffer(Display*, XID);
extern int  glXMakeContextCurrent(Display*, XID draw, XID read, GLXContext);
extern void glXDestroyContext(Display*, GLXContext);
extern glx_proc_t glXGetProcAddress(const unsigned char* name);
extern const unsigned char* glGetString(unsigned name);

enum {
    GLX_DOUBLEBUFFER = 5, GLX_RED_SIZE = 8, GLX_GREEN_SIZE = 9, GLX_BLUE_SIZE = 10, GLX_DEPTH_SIZE = 12,
    GLX_DRAWABLE_TYPE = 0x8010, GLX_RENDER_TYPE = 0x8011, GLX_RGBA_BIT = 1, GLX_PBUFFER_BIT = 4,
    GLX_PBUFFER_HEIGHT = 0x8040, GLX_PBUFFER_WIDTH = 0x8041,
    GLX_CONTEXT_MAJOR_VERSION_ARB = 0x2091, GLX_CONTEXT_MINOR_VERSION_ARB = 0x2092,
    GLX_CONTEXT_FLAGS_ARB = 0x2094, GLX_CONTEXT_FORWARD_COMPATIBLE_BIT_ARB = 2,
    GLX_CONTEXT_PROFILE_MASK_ARB = 0x9126, GLX_CONTEXT_CORE_PROFILE_BIT_ARB = 1,
};

static int g_major = 3, g_minor = 3, g_core = 0, g_forward = 0, g_w = 300, g_h = 300;
static Display*   g_dpy = NULL;
static GLXContext g_ctx = NULL;
static XID        g_pbuffer = 0;
static GLXFBConfig g_config = NULL;

static void crash_report(int sig)
{
    void* frames[64];
    int n = backtrace(frames, 64);
    static const char msg[] = "glut_glx: fatal signal inside the GL driver; backtrace:\n";
    if(write(2, msg, sizeof(msg) - 1)) {}
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}

void glutInitContextFlags(int flags)        { g_forward = (flags & GLUT_FORWARD_COMPATIBLE) != 0; }
void glutInitContextVersion(int major, int minor) { g_major = major; g_minor = minor; }
void glutInitContextProfile(int profile)    { g_core = (profile & GLUT_CORE_PROFILE) != 0; }
void glutInit(int* argc, char** argv)       { (void)argc; (void)argv; }
void glutInitDisplayMode(unsigned int mode) { (void)mode; }
void glutInitWindowSize(int w, int h)       { g_w = w; g_h = h; }

int glutCreateWindow(const char* title)
{
    (void)title;
    if(g_ctx) { fprintf(stderr, "glut_glx: one window at a time\n"); return 0; }
    if(getenv("GLUT_GLX_BACKTRACE")) { signal(SIGSEGV, crash_report); signal(SIGABRT, crash_report); }

    char version[32], glsl[32];
    snprintf(version, sizeof(version), "%d.%d%s", g_major, g_minor, g_core ? (g_forward ? "FC" : "") : "COMPAT");
    snprintf(glsl, sizeof(glsl), "%d%d0", g_major, g_minor);
    setenv("MESA_GL_VERSION_OVERRIDE", version, 0);
    setenv("MESA_GLSL_VERSION_OVERRIDE", glsl, 0);

    g_dpy = XOpenDisplay(NULL);
    const int config_attribs[] = { GLX_DRAWABLE_TYPE, GLX_PBUFFER_BIT, GLX_RENDER_TYPE, GLX_RGBA_BIT,
                                   GLX_RED_SIZE, 8, GLX_GREEN_SIZE, 8, GLX_BLUE_SIZE, 8, GLX_DEPTH_SIZE, 24, 0 };
    int n = 0;
    GLXFBConfig* configs = glXChooseFBConfig(g_dpy, 0, config_attribs, &n);
    if(!configs || n < 1) { fprintf(stderr, "glut_glx: no framebuffer configuration\n"); return 0; }
    g_config = configs[0];
    XFree(configs);

    typedef GLXContext (*create_t)(Display*, GLXFBConfig, GLXContext, int, const int*);
    create_t create = (create_t)glXGetProcAddress((const unsigned char*)"glXCreateContextAttribsARB");
    if(!create) { fprintf(stderr, "glut_glx: no glXCreateContextAttribsARB\n"); return 0; }
    const int context_attribs[] = { GLX_CONTEXT_MAJOR_VERSION_ARB, g_major, GLX_CONTEXT_MINOR_VERSION_ARB, g_minor,
                                    GLX_CONTEXT_FLAGS_ARB, g_forward ? GLX_CONTEXT_FORWARD_COMPATIBLE_BIT_ARB : 0,
                                    GLX_CONTEXT_PROFILE_MASK_ARB, g_core ? GLX_CONTEXT_CORE_PROFILE_BIT_ARB : 2, 0 };
    g_ctx = create(g_dpy, g_config, NULL, 1, context_attribs);
    if(!g_ctx) { fprintf(stderr, "glut_glx: no %d.%d context\n", g_major, g_minor); return 0; }

    const int pbuffer_attribs[] = { GLX_PBUFFER_WIDTH, g_w, GLX_PBUFFER_HEIGHT, g_h, 0 };
    g_pbuffer = glXCreatePbuffer(g_dpy, g_config, pbuffer_attribs);
    if(!g_pbuffer || !glXMakeContextCurrent(g_dpy, g_pbuffer, g_pbuffer, g_ctx)) {
        fprintf(stderr, "glut_glx: cannot make the context current\n");
        return 0;
    }
    if(getenv("GLUT_GLX_VERBOSE"))
        fprintf(stderr, "glut_glx: GL_VERSION %s, GL_RENDERER %s\n", glGetString(0x1F02), glGetString(0x1F01));
    return 1;
}

void glutHideWindow(void) {}
int  glutExtensionSupported(const char* ext)
{
    /* the four the reference asks about (vertex/fragment shaders, VBOs, FBOs) are core since GL 2.0/3.0; a core
       profile does not list them, as on any modern driver where the reference runs */
    (void)ext;
    return g_major >= 3;
}
void glutSetWindow(int id) { (void)id; if(g_ctx) glXMakeContextCurrent(g_dpy, g_pbuffer, g_pbuffer, g_ctx); }
void glutDestroyWindow(int id)
{
    (void)id;
    if(!g_ctx) return;
    glXMakeContextCurrent(g_dpy, 0, 0, NULL);
    glXDestroyPbuffer(g_dpy, g_pbuffer);
    glXDestroyContext(g_dpy, g_ctx);
    g_ctx = NULL; g_pbuffer = 0;
}
void glutExit(void) {}

/* which GL is underneath (for reports) */
const char* glut_glx_renderer(void) { return g_ctx ? (const char*)glGetString(0x1F01) : ""; }
const char* glut_glx_version(void)  { return g_ctx ? (const char*)glGetString(0x1F02) : ""; }
