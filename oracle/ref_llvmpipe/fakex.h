// Minimal hand-declared Xlib structures (x86-64 layouts) for a display-less stub.
#pragma once
#include <stddef.h>
typedef unsigned long XID; typedef XID Window, Drawable, Pixmap, Colormap, Font, VisualID; typedef char* XPointer; typedef int Bool; typedef int Status;
typedef struct _XExtData XExtData;
typedef struct { XExtData* ext_data; VisualID visualid; int c_class; unsigned long red_mask, green_mask, blue_mask; int bits_per_rgb; int map_entries; } Visual;
typedef struct { int depth; int nvisuals; Visual* visuals; } Depth;
typedef struct _XGC* GC;
struct _XDisplay;
typedef struct { XExtData* ext_data; struct _XDisplay* display; Window root; int width, height, mwidth, mheight; int ndepths; Depth* depths; int root_depth; Visual* root_visual; GC default_gc; Colormap cmap; unsigned long white_pixel, black_pixel; int max_maps, min_maps; int backing_store; Bool save_unders; long root_input_mask; } Screen;
typedef struct { XExtData* ext_data; int depth; int bits_per_pixel; int scanline_pad; } ScreenFormat;
typedef struct { int extension, major_opcode, first_event, first_error; } XExtCodes;
typedef struct _XExten { struct _XExten* next; XExtCodes codes; void* create_GC,*copy_GC,*flush_GC,*free_GC,*create_Font,*free_Font,*close_display,*error,*error_string; char* name; void* error_values; void* before_flush; struct _XExten* next_flush; } _XExtension;
typedef struct _XDisplay {
  XExtData* ext_data; void* free_funcs; int fd; int conn_checker; int proto_major_version, proto_minor_version; char* vendor;
  XID resource_base, resource_mask, resource_id; int resource_shift; XID (*resource_alloc)(struct _XDisplay*);
  int byte_order, bitmap_unit, bitmap_pad, bitmap_bit_order; int nformats; ScreenFormat* pixmap_format; int vnumber; int release;
  void *head,*tail; int qlen; unsigned long last_request_read, request; char *last_req,*buffer,*bufptr,*bufmax; unsigned max_request_size;
  void* db; int (*synchandler)(struct _XDisplay*); char* display_name; int default_screen; int nscreens; Screen* screens;
  unsigned long motion_buffer; volatile unsigned long flags; int min_keycode, max_keycode; void* keysyms; void* modifiermap; int keysyms_per_keycode;
  char* xdefaults; char* scratch_buffer; unsigned long scratch_length; int ext_number; _XExtension* ext_procs;
  char tail_pad[8192];   /* event_vec, wire_vec, lock_fns ... all NULL */
} Display;
typedef struct { Visual* visual; VisualID visualid; int screen; int depth; int c_class; unsigned long red_mask, green_mask, blue_mask; int colormap_size; int bits_per_rgb; } XVisualInfo;
typedef struct _XImage { int width, height, xoffset, format; char* data; int byte_order, bitmap_unit, bitmap_bit_order, bitmap_pad, depth, bytes_per_line, bits_per_pixel; unsigned long red_mask, green_mask, blue_mask; XPointer obdata;
  struct { struct _XImage* (*create_image)(); int (*destroy_image)(struct _XImage*); unsigned long (*get_pixel)(struct _XImage*,int,int); int (*put_pixel)(struct _XImage*,int,int,unsigned long); struct _XImage* (*sub_image)(); int (*add_pixel)(struct _XImage*,long);} f; } XImage;
Display* FakeOpenDisplay(void);
