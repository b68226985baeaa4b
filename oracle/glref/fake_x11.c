// A do-nothing Xlib: the 27 entry points (and the structures its macros read) that the Mesa 18.1 llvmpipe libGL bundled with
// Nsight Compute needs to create an off-screen GLX context with no X server. TEST INFRASTRUCTURE ONLY (oracle/glref/README):
// built into oracle/_ref/glx/libX11.so.6 + libXext.so.6 by oracle/build_ref.py gl; rendering goes to framebuffer objects, so
// no pixel ever reaches an X drawable and every drawing request below is a no-op.
#include <stdlib.h>
#include <string.h>
#include "fake_x11.h"

static Visual g_visual = {0, 0x21, 4 /* TrueColor */, 0xff0000ul, 0x00ff00ul, 0x0000fful, 8, 256};
static Depth g_depth = {24, 1, &g_visual};
static Screen g_screen;
static struct _XDisplay g_display;
static struct _XGC { int dummy; } g_gc;
static ScreenFormat g_format = {0, 24, 32, 32};

Display* XOpenDisplay(const char* name) {
    (void)name;
    memset(&g_screen, 0, sizeof g_screen);
    memset(&g_display, 0, sizeof g_display);
    g_screen.display = &g_display;
    g_screen.root = 0x100;
    g_screen.width = 1024; g_screen.height = 768; g_screen.mwidth = 270; g_screen.mheight = 203;
    g_screen.ndepths = 1; g_screen.depths = &g_depth;
    g_screen.root_depth = 24; g_screen.root_visual = &g_visual; g_screen.default_gc = (GC)&g_gc; g_screen.cmap = 0x20;
    g_screen.white_pixel = 0xffffff; g_screen.black_pixel = 0;
    g_display.vendor = (char*)"stillleben-b200 fake X";
    g_display.proto_major_version = 11;
    g_display.byte_order = 0 /* LSBFirst */; g_display.bitmap_unit = 32; g_display.bitmap_pad = 32; g_display.bitmap_bit_order = 0;
    g_display.nformats = 1; g_display.pixmap_format = &g_format;
    g_display.display_name = (char*)":fake";
    g_display.default_screen = 0; g_display.nscreens = 1; g_display.screens = &g_screen;
    return &g_display;
}
int XCloseDisplay(Display* d) { (void)d; return 0; }

static int img_destroy(XImage* im) { if (im) { free(im->data); free(im); } return 1; }
static unsigned long img_get(XImage* im, int x, int y) { (void)im; (void)x; (void)y; return 0; }
static int img_put(XImage* im, int x, int y, unsigned long p) { (void)im; (void)x; (void)y; (void)p; return 0; }
static int img_add(XImage* im, long v) { (void)im; (void)v; return 0; }
static XImage* img_sub(XImage* im, int x, int y, unsigned w, unsigned h) { (void)im; (void)x; (void)y; (void)w; (void)h; return 0; }

XImage* XCreateImage(Display* d, Visual* v, unsigned depth, int format, int offset, char* data, unsigned w, unsigned h, int pad, int bpl) {
    (void)d;
    XImage* im = (XImage*)calloc(1, sizeof(XImage));
    im->width = (int)w; im->height = (int)h; im->xoffset = offset; im->format = format; im->data = data;
    im->byte_order = 0; im->bitmap_unit = 32; im->bitmap_bit_order = 0; im->bitmap_pad = pad; im->depth = (int)depth;
    im->bits_per_pixel = depth > 16 ? 32 : depth > 8 ? 16 : 8;
    im->bytes_per_line = bpl ? bpl : (int)(((w * (unsigned)im->bits_per_pixel + (unsigned)pad - 1) / (unsigned)pad) * (unsigned)pad / 8);
    if (v) { im->red_mask = v->red_mask; im->green_mask = v->green_mask; im->blue_mask = v->blue_mask; }
    im->f.create_image = 0; im->f.destroy_image = img_destroy; im->f.get_pixel = img_get; im->f.put_pixel = img_put;
    im->f.sub_image = img_sub; im->f.add_pixel = img_add;
    return im;
}
XImage* XGetImage(Display* d, Drawable dr, int x, int y, unsigned w, unsigned h, unsigned long mask, int format) {
    (void)dr; (void)x; (void)y; (void)mask;
    return XCreateImage(d, &g_visual, 24, format, 0, (char*)calloc((size_t)w * h, 4), w, h, 32, 0);
}
XVisualInfo* XGetVisualInfo(Display* d, long mask, XVisualInfo* tmpl, int* n) {
    (void)d;
    *n = 0;
    if ((mask & 0x1 /* VisualIDMask */) && tmpl->visualid != g_visual.visualid) return 0;
    if ((mask & 0x2 /* VisualScreenMask */) && tmpl->screen != 0) return 0;
    if ((mask & 0x4 /* VisualDepthMask */) && tmpl->depth != 24) return 0;
    if ((mask & 0x8 /* VisualClassMask */) && tmpl->c_class != 4) return 0;
    XVisualInfo* vi = (XVisualInfo*)calloc(1, sizeof(XVisualInfo));
    vi->visual = &g_visual; vi->visualid = g_visual.visualid; vi->screen = 0; vi->depth = 24; vi->c_class = 4;
    vi->red_mask = g_visual.red_mask; vi->green_mask = g_visual.green_mask; vi->blue_mask = g_visual.blue_mask;
    vi->colormap_size = 256; vi->bits_per_rgb = 8;
    *n = 1;
    return vi;
}
int XGetGeometry(Display* d, Drawable dr, Window* root, int* x, int* y, unsigned* w, unsigned* h, unsigned* bw, unsigned* depth) {
    (void)d; (void)dr;
    if (root) *root = g_screen.root;
    if (x) *x = 0;
    if (y) *y = 0;
    if (w) *w = 64;
    if (h) *h = 64;
    if (bw) *bw = 0;
    if (depth) *depth = 24;
    return 1;
}
int XGetWindowAttributes(Display* d, Window w, XWindowAttributes* a) {
    (void)d; (void)w;
    memset(a, 0, sizeof *a);
    a->width = 64; a->height = 64; a->depth = 24; a->visual = &g_visual; a->root = g_screen.root; a->c_class = 1 /* InputOutput */;
    a->colormap = g_screen.cmap; a->map_installed = 1; a->map_state = 2 /* IsViewable */; a->screen = &g_screen;
    return 1;
}
XExtCodes* XAddExtension(Display* d) {   // new extensions go to the head of the display's list (Mesa reads d->ext_procs right after)
    _XExtension* e = (_XExtension*)calloc(1, sizeof(_XExtension));
    e->codes.extension = d->ext_number++;
    e->next = d->ext_procs;
    d->ext_procs = e;
    return &e->codes;
}
Colormap XCreateColormap(Display* d, Window w, Visual* v, int alloc) { (void)d; (void)w; (void)v; (void)alloc; return 0x21; }
GC XCreateGC(Display* d, Drawable dr, unsigned long mask, void* values) { (void)d; (void)dr; (void)mask; (void)values; return (GC)calloc(1, 64); }
int XFreeGC(Display* d, GC gc) { (void)d; free(gc); return 1; }
Pixmap XCreatePixmap(Display* d, Drawable dr, unsigned w, unsigned h, unsigned depth) { (void)d; (void)dr; (void)w; (void)h; (void)depth; static Pixmap next = 0x400; return next++; }
int XFreePixmap(Display* d, Pixmap p) { (void)d; (void)p; return 1; }
int XDrawString16(Display* d, Drawable dr, GC gc, int x, int y, const void* s, int n) { (void)d; (void)dr; (void)gc; (void)x; (void)y; (void)s; (void)n; return 0; }
int XFillRectangle(Display* d, Drawable dr, GC gc, int x, int y, unsigned w, unsigned h) { (void)d; (void)dr; (void)gc; (void)x; (void)y; (void)w; (void)h; return 1; }
int XFlush(Display* d) { (void)d; return 1; }
int XFree(void* p) { free(p); return 1; }
int XFreeFontInfo(char** names, void* info, int n) { (void)names; (void)info; (void)n; return 1; }
int XPutImage(Display* d, Drawable dr, GC gc, XImage* im, int sx, int sy, int dx, int dy, unsigned w, unsigned h) { (void)d; (void)dr; (void)gc; (void)im; (void)sx; (void)sy; (void)dx; (void)dy; (void)w; (void)h; return 0; }
int XQueryExtension(Display* d, const char* name, int* op, int* ev, int* err) { (void)d; (void)name; if (op) *op = 0; if (ev) *ev = 0; if (err) *err = 0; return 0; }
void* XQueryFont(Display* d, XID id) { (void)d; (void)id; return 0; }
typedef int (*XErrorHandler)(Display*, void*);
XErrorHandler XSetErrorHandler(XErrorHandler h) { static XErrorHandler cur; XErrorHandler old = cur; cur = h; return old; }
int XSetForeground(Display* d, GC gc, unsigned long fg) { (void)d; (void)gc; (void)fg; return 1; }
int XSetFunction(Display* d, GC gc, int f) { (void)d; (void)gc; (void)f; return 1; }
int XSync(Display* d, int discard) { (void)d; (void)discard; return 1; }
typedef int (*XSyncProc)(Display*);
XSyncProc XSynchronize(Display* d, int onoff) { (void)d; (void)onoff; return 0; }
// libXext (MIT-SHM): the extension query above says "absent", so these are never reached
int XShmAttach(Display* d, void* info) { (void)d; (void)info; return 0; }
XImage* XShmCreateImage(Display* d, Visual* v, unsigned depth, int format, char* data, void* info, unsigned w, unsigned h) { (void)d; (void)v; (void)depth; (void)format; (void)data; (void)info; (void)w; (void)h; return 0; }
int XShmPutImage(Display* d, Drawable dr, GC gc, XImage* im, int sx, int sy, int dx, int dy, unsigned w, unsigned h, int ev) { (void)d; (void)dr; (void)gc; (void)im; (void)sx; (void)sy; (void)dx; (void)dy; (void)w; (void)h; (void)ev; return 0; }
// Xlib's global lock hooks (Xlibint.h: _XLockMutex(lock) calls *_XLockMutex_fn when it is non-null): left null = unthreaded Xlib
void (*_XLockMutex_fn)(void*) = 0;
void (*_XUnlockMutex_fn)(void*) = 0;
void* _Xglobal_lock = 0;
