// Xlib structure layouts (X11R6 ABI, LP64) as far as Mesa's xlib GLX front end and this harness read them. See fake_x11.c.
#pragma once
typedef unsigned long XID, Window, Drawable, Pixmap, Colormap, VisualID;
typedef char* XPointer;
typedef struct _XGC* GC;
typedef struct { void* ext_data; VisualID visualid; int c_class; unsigned long red_mask, green_mask, blue_mask; int bits_per_rgb; int map_entries; } Visual;
typedef struct { int depth; int nvisuals; Visual* visuals; } Depth;
struct _XDisplay;
typedef struct {
    void* ext_data; struct _XDisplay* display; Window root; int width, height; int mwidth, mheight; int ndepths; Depth* depths;
    int root_depth; Visual* root_visual; GC default_gc; Colormap cmap; unsigned long white_pixel, black_pixel; int max_maps, min_maps;
    int backing_store; int save_unders; long root_input_mask;
} Screen;
typedef struct { void* ext_data; int depth; int bits_per_pixel; int scanline_pad; } ScreenFormat;
typedef struct _XDisplay {
    void* ext_data; void* private1; int fd; int private2; int proto_major_version; int proto_minor_version; char* vendor;
    XID private3, private4, private5; int private6; XID (*resource_alloc)(struct _XDisplay*); int byte_order; int bitmap_unit; int bitmap_pad;
    int bitmap_bit_order; int nformats; ScreenFormat* pixmap_format; int private8; int release; void *private9, *private10; int qlen;
    unsigned long last_request_read; unsigned long request; XPointer private11, private12, private13, private14; unsigned max_request_size;
    void* db; int (*private15)(struct _XDisplay*); char* display_name; int default_screen; int nscreens; Screen* screens;
    unsigned long motion_buffer; unsigned long private16; int min_keycode, max_keycode; XPointer private17, private18; int private19;
    char* xdefaults;
    // Xlib's private continuation (Xlibint.h): Mesa's GLX front end hangs its close-display hook on the head of ext_procs
    char* scratch_buffer; unsigned long scratch_length; int ext_number; struct _XExten* ext_procs;
    char slack[1024];
} Display;
typedef struct { Visual* visual; VisualID visualid; int screen; int depth; int c_class; unsigned long red_mask, green_mask, blue_mask; int colormap_size; int bits_per_rgb; } XVisualInfo;
typedef struct _XImage {
    int width, height; int xoffset; int format; char* data; int byte_order; int bitmap_unit; int bitmap_bit_order; int bitmap_pad; int depth;
    int bytes_per_line; int bits_per_pixel; unsigned long red_mask, green_mask, blue_mask; XPointer obdata;
    struct funcs {
        struct _XImage* (*create_image)(void);
        int (*destroy_image)(struct _XImage*);
        unsigned long (*get_pixel)(struct _XImage*, int, int);
        int (*put_pixel)(struct _XImage*, int, int, unsigned long);
        struct _XImage* (*sub_image)(struct _XImage*, int, int, unsigned int, unsigned int);
        int (*add_pixel)(struct _XImage*, long);
    } f;
} XImage;
typedef struct {
    int x, y; int width, height; int border_width; int depth; Visual* visual; Window root; int c_class; int bit_gravity; int win_gravity;
    int backing_store; unsigned long backing_planes; unsigned long backing_pixel; int save_under; Colormap colormap; int map_installed;
    int map_state; long all_event_masks; long your_event_mask; long do_not_propagate_mask; int override_redirect; Screen* screen;
} XWindowAttributes;
typedef struct { int extension; int major_opcode; int first_event; int first_error; } XExtCodes;
typedef struct _XExten {   // Xlibint.h _XExtension: list link, the public codes, then fourteen hook / name slots
    struct _XExten* next; XExtCodes codes; void* hooks[9]; char* name; void* more[3];
} _XExtension;
#ifdef __cplusplus
extern "C" {
#endif
Display* XOpenDisplay(const char* name);
int XCloseDisplay(Display*);
int XFree(void*);
#ifdef __cplusplus
}
#endif
