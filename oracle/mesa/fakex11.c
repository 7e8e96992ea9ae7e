/* fakex11.c -- TEST INFRASTRUCTURE (oracle).
 *
 * A display-less stand-in for libX11 / libXext, just big enough for the Xlib flavour of Mesa's libGL
 * (gallium "xlib" winsys with the llvmpipe software rasteriser) to create a context and a pbuffer.  The image
 * holds such a libGL (shipped with Nsight Compute for its own UI) but neither libX11 nor an X server, so this
 * file supplies the ~30 Xlib entry points that libGL imports.  Nothing is ever displayed: the reference
 * renders into a framebuffer object and reads it back with glReadPixels(), so the only things the "server"
 * has to answer are "which visuals exist" (one: 24-bit TrueColor) and "how big is this drawable".
 *
 * Built twice by oracle/Makefile, as oracle/_ref/libX11.so.6 and oracle/_ref/libXext.so.6 (the sonames libGL
 * asks for).  Structure layouts follow the public X11 headers (Xlib.h, Xutil.h) and, for the two private
 * fields Mesa's GLX front end touches (Display::ext_procs and the extension record), Xlibint.h; none of those
 * headers exist in the image, so they are restated here and the offsets Mesa uses were confirmed from its
 * disassembly (ext_procs at 0x140, record name at 0x60, close_display at 0x48, sizeof(XVisualInfo) == 0x40).
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned long XID;
typedef char* XPointer;

typedef struct {
    void* ext_data;
    XID visualid;
    int c_class;
    unsigned long red_mask, green_mask, blue_mask;
    int bits_per_rgb;
    int map_entries;
} Visual;

typedef struct {
    Visual* visual;
    XID visualid;
    int screen;
    int depth;
    int c_class;
    unsigned long red_mask, green_mask, blue_mask;
    int colormap_size;
    int bits_per_rgb;
} XVisualInfo;

struct FakeDisplay;

typedef struct {
    void* ext_data;
    struct FakeDisplay* display;
    XID root;
    int width, height;
    int mwidth, mheight;
    int ndepths;
    void* depths;
    int root_depth;
    Visual* root_visual;
    void* default_gc;
    XID cmap;
    unsigned long white_pixel, black_pixel;
    int max_maps, min_maps;
    int backing_store;
    int save_unders;
    long root_input_mask;
} Screen;

typedef struct { int extension, major_opcode, first_event, first_error; } XExtCodes;

typedef struct FakeExtension {      /* Xlibint.h: struct _XExten */
    struct FakeExtension* next;
    XExtCodes codes;
    void* hooks_before_name[9];     /* create_GC .. error_string; close_display is the 7th (offset 0x48) */
    char* name;                     /* offset 0x60 */
    void* more[16];
} FakeExtension;

typedef struct FakeDisplay {        /* Xlibint.h: struct _XDisplay, up to the last field Mesa reads */
    void* ext_data;
    void* free_funcs;
    int fd;
    int conn_checker;
    int proto_major_version, proto_minor_version;
    char* vendor;
    XID resource_base, resource_mask, resource_id;
    int resource_shift;
    XID (*resource_alloc)(struct FakeDisplay*);
    int byte_order, bitmap_unit, bitmap_pad, bitmap_bit_order;
    int nformats;
    void* pixmap_format;
    int vnumber;
    int release;
    void *head, *tail;
    int qlen;
    unsigned long last_request_read, request;
    char *last_req, *buffer, *bufptr, *bufmax;
    unsigned max_request_size;
    void* db;
    int (*synchandler)(struct FakeDisplay*);
    char* display_name;
    int default_screen;
    int nscreens;
    Screen* screens;
    unsigned long motion_buffer;
    unsigned long flags;
    int min_keycode, max_keycode;
    void* keysyms;
    void* modifiermap;
    int keysyms_per_keycode;
    char* xdefaults;
    char* scratch_buffer;
    unsigned long scratch_length;
    int ext_number;
    FakeExtension* ext_procs;
    char rest_of_xlib_private_state[2048];   /* zero: no lock functions, no hooks */
} Display;

_Static_assert(sizeof(XVisualInfo) == 0x40, "XVisualInfo layout");
_Static_assert(offsetof(Display, default_screen) == 224 && offsetof(Display, screens) == 232, "Display layout");
_Static_assert(offsetof(Display, ext_procs) == 0x140, "Display layout (private part)");
_Static_assert(offsetof(FakeExtension, name) == 0x60 && offsetof(FakeExtension, hooks_before_name[6]) == 0x48,
               "extension record layout");
_Static_assert(offsetof(Screen, root) == 16 && offsetof(Screen, root_depth) == 56 && offsetof(Screen, root_visual) == 64,
               "Screen layout");

typedef struct XImage {
    int width, height;
    int xoffset;
    int format;
    char* data;
    int byte_order;
    int bitmap_unit;
    int bitmap_bit_order;
    int bitmap_pad;
    int depth;
    int bytes_per_line;
    int bits_per_pixel;
    unsigned long red_mask, green_mask, blue_mask;
    XPointer obdata;
    struct {
        struct XImage* (*create_image)(void);
        int (*destroy_image)(struct XImage*);
        unsigned long (*get_pixel)(struct XImage*, int, int);
        int (*put_pixel)(struct XImage*, int, int, unsigned long);
        struct XImage* (*sub_image)(struct XImage*, int, int, unsigned, unsigned);
        int (*add_pixel)(struct XImage*, long);
    } f;
} XImage;

enum { TrueColor = 4, ZPixmap = 2, LSBFirst = 0 };
enum { VisualIDMask = 0x1, VisualScreenMask = 0x2, VisualDepthMask = 0x4, VisualClassMask = 0x8 };

/* ---- the one screen, the one visual ---- */
static Visual  g_visual = { NULL, 0x21, TrueColor, 0xff0000, 0x00ff00, 0x0000ff, 8, 256 };
static Screen  g_screen;
static Display g_display;
static int     g_open = 0;

/* pixmaps behind pbuffers; Mesa does not give them all back, so the table grows */
typedef struct { XID id; unsigned w, h, depth; } drawable_t;
static drawable_t* g_drawables = NULL;
static int g_ndrawables = 0;
static XID g_next_id = 0x400001;

/* Xlib's locking hooks (data symbols libGL imports); NULL = single-threaded Xlib */
void (*_XLockMutex_fn)(void*) = NULL;
void (*_XUnlockMutex_fn)(void*) = NULL;
void* _Xglobal_lock = NULL;

Display* XOpenDisplay(const char* name)
{
    (void)name;
    if(!g_open) {
        memset(&g_display, 0, sizeof(g_display));
        memset(&g_screen, 0, sizeof(g_screen));
        g_screen.display = &g_display;
        g_screen.root = 0x100;
        g_screen.width = 4096; g_screen.height = 4096; g_screen.mwidth = 1084; g_screen.mheight = 1084;
        g_screen.root_depth = 24;
        g_screen.root_visual = &g_visual;
        g_screen.cmap = 0x20;
        g_screen.white_pixel = 0xffffff;
        g_display.fd = -1;
        g_display.proto_major_version = 11;
        g_display.vendor = "fakex11 (oracle test infrastructure; no server)";
        g_display.byte_order = LSBFirst; g_display.bitmap_unit = 32; g_display.bitmap_pad = 32;
        g_display.display_name = ":fake";
        g_display.nscreens = 1;
        g_display.screens = &g_screen;
        g_open = 1;
    }
    return &g_display;
}
int XCloseDisplay(Display* dpy) { (void)dpy; return 0; }

XExtCodes* XAddExtension(Display* dpy)
{
    FakeExtension* e = calloc(1, sizeof(*e));
    e->codes.extension = dpy->ext_number++;
    e->next = dpy->ext_procs;
    dpy->ext_procs = e;
    return &e->codes;
}

XVisualInfo* XGetVisualInfo(Display* dpy, long mask, XVisualInfo* tmpl, int* nitems)
{
    (void)dpy;
    *nitems = 0;
    if((mask & VisualIDMask)     && tmpl->visualid != g_visual.visualid) return NULL;
    if((mask & VisualScreenMask) && tmpl->screen   != 0)                 return NULL;
    if((mask & VisualDepthMask)  && tmpl->depth    != 24)                return NULL;
    if((mask & VisualClassMask)  && tmpl->c_class  != TrueColor)         return NULL;
    if(mask & ~(long)(VisualIDMask | VisualScreenMask | VisualDepthMask | VisualClassMask)) return NULL;
    XVisualInfo* v = calloc(1, sizeof(*v));
    v->visual = &g_visual; v->visualid = g_visual.visualid; v->screen = 0; v->depth = 24; v->c_class = TrueColor;
    v->red_mask = g_visual.red_mask; v->green_mask = g_visual.green_mask; v->blue_mask = g_visual.blue_mask;
    v->colormap_size = 256; v->bits_per_rgb = 8;
    *nitems = 1;
    return v;
}
int XFree(void* p) { free(p); return 1; }

static int destroy_image(XImage* img) { if(img) { free(img->data); free(img); } return 1; }
XImage* XCreateImage(Display* dpy, Visual* visual, unsigned depth, int format, int offset, char* data,
                     unsigned width, unsigned height, int bitmap_pad, int bytes_per_line)
{
    (void)dpy;
    XImage* img = calloc(1, sizeof(*img));
    img->width = (int)width; img->height = (int)height; img->xoffset = offset; img->format = format; img->data = data;
    img->byte_order = LSBFirst; img->bitmap_unit = 32; img->bitmap_bit_order = LSBFirst; img->bitmap_pad = bitmap_pad;
    img->depth = (int)depth;
    img->bits_per_pixel = depth <= 1 ? 1 : depth <= 8 ? 8 : depth <= 16 ? 16 : 32;
    img->bytes_per_line = bytes_per_line ? bytes_per_line : (int)(((size_t)width * img->bits_per_pixel + 31) / 32 * 4);
    if(visual) { img->red_mask = visual->red_mask; img->green_mask = visual->green_mask; img->blue_mask = visual->blue_mask; }
    img->f.destroy_image = destroy_image;
    return img;
}
XImage* XGetImage(Display* dpy, XID d, int x, int y, unsigned w, unsigned h, unsigned long planes, int format)
{ (void)dpy; (void)d; (void)x; (void)y; (void)w; (void)h; (void)planes; (void)format; return NULL; }
int XPutImage(Display* dpy, XID d, void* gc, XImage* img, int sx, int sy, int dx, int dy, unsigned w, unsigned h)
{ (void)dpy; (void)d; (void)gc; (void)img; (void)sx; (void)sy; (void)dx; (void)dy; (void)w; (void)h; return 0; }

XID XCreatePixmap(Display* dpy, XID d, unsigned w, unsigned h, unsigned depth)
{
    (void)dpy; (void)d;
    int k = 0;
    while(k < g_ndrawables && g_drawables[k].id) k++;
    if(k == g_ndrawables) {
        int n = g_ndrawables ? 2 * g_ndrawables : 64;
        drawable_t* grown = realloc(g_drawables, (size_t)n * sizeof(*grown));
        if(!grown) { fprintf(stderr, "fakex11: out of memory\n"); return 0; }
        memset(grown + g_ndrawables, 0, (size_t)(n - g_ndrawables) * sizeof(*grown));
        g_drawables = grown; g_ndrawables = n;
    }
    g_drawables[k].id = g_next_id++; g_drawables[k].w = w; g_drawables[k].h = h; g_drawables[k].depth = depth;
    return g_drawables[k].id;
}
int XFreePixmap(Display* dpy, XID p)
{
    (void)dpy;
    for(int k = 0; k < g_ndrawables; k++) if(g_drawables[k].id == p) g_drawables[k].id = 0;
    return 1;
}
int XGetGeometry(Display* dpy, XID d, XID* root, int* x, int* y, unsigned* w, unsigned* h, unsigned* bw, unsigned* depth)
{
    (void)dpy;
    *root = g_screen.root; *x = *y = 0; *bw = 0;
    for(int k = 0; k < g_ndrawables; k++)
        if(g_drawables[k].id == d) { *w = g_drawables[k].w; *h = g_drawables[k].h; *depth = g_drawables[k].depth; return 1; }
    *w = (unsigned)g_screen.width; *h = (unsigned)g_screen.height; *depth = 24;
    return d == g_screen.root;
}

typedef struct {
    int x, y, width, height, border_width, depth;
    Visual* visual;
    XID root;
    int c_class, bit_gravity, win_gravity, backing_store;
    unsigned long backing_planes, backing_pixel;
    int save_under;
    XID colormap;
    int map_installed, map_state;
    long all_event_masks, your_event_mask, do_not_propagate_mask;
    int override_redirect;
    Screen* screen;
} XWindowAttributes;
int XGetWindowAttributes(Display* dpy, XID w, XWindowAttributes* a)
{
    XID root; int x, y; unsigned ww, hh, bw, depth;
    memset(a, 0, sizeof(*a));
    if(!XGetGeometry(dpy, w, &root, &x, &y, &ww, &hh, &bw, &depth)) return 0;
    a->width = (int)ww; a->height = (int)hh; a->depth = (int)depth; a->visual = &g_visual; a->root = root;
    a->c_class = 1; a->colormap = g_screen.cmap; a->map_state = 2; a->screen = &g_screen;
    return 1;
}

static int g_gc_storage[32];
void* XCreateGC(Display* dpy, XID d, unsigned long mask, void* values) { (void)dpy; (void)d; (void)mask; (void)values; return g_gc_storage; }
int XFreeGC(Display* dpy, void* gc) { (void)dpy; (void)gc; return 1; }
int XSetFunction(Display* dpy, void* gc, int f) { (void)dpy; (void)gc; (void)f; return 1; }
int XSetForeground(Display* dpy, void* gc, unsigned long c) { (void)dpy; (void)gc; (void)c; return 1; }
int XFillRectangle(Display* dpy, XID d, void* gc, int x, int y, unsigned w, unsigned h)
{ (void)dpy; (void)d; (void)gc; (void)x; (void)y; (void)w; (void)h; return 1; }
int XDrawString16(Display* dpy, XID d, void* gc, int x, int y, const void* s, int n)
{ (void)dpy; (void)d; (void)gc; (void)x; (void)y; (void)s; (void)n; return 0; }
int XFlush(Display* dpy) { (void)dpy; return 1; }
int XSync(Display* dpy, int discard) { (void)dpy; (void)discard; return 1; }
void* XSynchronize(Display* dpy, int onoff) { (void)dpy; (void)onoff; return NULL; }
typedef int (*XErrorHandler)(Display*, void*);
static XErrorHandler g_handler = NULL;
XErrorHandler XSetErrorHandler(XErrorHandler h) { XErrorHandler old = g_handler; g_handler = h; return old; }
int XQueryExtension(Display* dpy, const char* name, int* major, int* event, int* error)
{ (void)dpy; (void)name; *major = *event = *error = 0; return 0; }
XID XCreateColormap(Display* dpy, XID w, Visual* v, int alloc) { (void)dpy; (void)w; (void)v; (void)alloc; return g_screen.cmap; }
void* XQueryFont(Display* dpy, XID id) { (void)dpy; (void)id; return NULL; }
int XFreeFontInfo(char** names, void* info, int n) { (void)names; (void)info; (void)n; return 1; }

/* ---- libXext (MIT-SHM): never available ---- */
int XShmQueryExtension(Display* dpy) { (void)dpy; return 0; }
int XShmAttach(Display* dpy, void* info) { (void)dpy; (void)info; return 0; }
int XShmDetach(Display* dpy, void* info) { (void)dpy; (void)info; return 0; }
XImage* XShmCreateImage(Display* dpy, Visual* v, unsigned depth, int format, char* data, void* info, unsigned w, unsigned h)
{ (void)dpy; (void)v; (void)depth; (void)format; (void)data; (void)info; (void)w; (void)h; return NULL; }
int XShmPutImage(Display* dpy, XID d, void* gc, XImage* img, int sx, int sy, int dx, int dy, unsigned w, unsigned h, int ev)
{ (void)dpy; (void)d; (void)gc; (void)img; (void)sx; (void)sy; (void)dx; (void)dy; (void)w; (void)h; (void)ev; return 0; }
