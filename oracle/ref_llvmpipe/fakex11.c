#include "fakex.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#define LOG(...) do{ if(getenv("FAKEX_DEBUG")) fprintf(stderr,"[fakex] " __VA_ARGS__);}while(0)
void (*_XLockMutex_fn)(void*) = 0; void (*_XUnlockMutex_fn)(void*) = 0; void* _Xglobal_lock = 0;
static Visual g_vis; static Depth g_depth; static Screen g_scr; static ScreenFormat g_fmt[2]; static struct _XGC{int x;} g_gc;
typedef struct { XID id; unsigned w,h,depth; } Pix; static Pix g_pix[4096]; static int g_npix; static XID g_next=0x400001;
Display* FakeOpenDisplay(void){ Display* d=calloc(1,sizeof(Display));
  g_vis.visualid=0x21; g_vis.c_class=4/*TrueColor*/; g_vis.red_mask=0xff0000; g_vis.green_mask=0xff00; g_vis.blue_mask=0xff; g_vis.bits_per_rgb=8; g_vis.map_entries=256;
  g_depth.depth=24; g_depth.nvisuals=1; g_depth.visuals=&g_vis;
  g_scr.display=d; g_scr.root=0x100; g_scr.width=1024; g_scr.height=768; g_scr.mwidth=300; g_scr.mheight=200; g_scr.ndepths=1; g_scr.depths=&g_depth; g_scr.root_depth=24; g_scr.root_visual=&g_vis; g_scr.default_gc=&g_gc; g_scr.cmap=0x20; g_scr.white_pixel=0xffffff; g_scr.max_maps=1; g_scr.min_maps=1;
  g_fmt[0].depth=24; g_fmt[0].bits_per_pixel=32; g_fmt[0].scanline_pad=32; g_fmt[1].depth=32; g_fmt[1].bits_per_pixel=32; g_fmt[1].scanline_pad=32;
  d->fd=-1; d->proto_major_version=11; d->vendor="fakex"; d->byte_order=0; d->bitmap_unit=32; d->bitmap_pad=32; d->bitmap_bit_order=0; d->nformats=2; d->pixmap_format=g_fmt; d->vnumber=11; d->release=1;
  d->display_name=":fake"; d->default_screen=0; d->nscreens=1; d->screens=&g_scr; d->max_request_size=65535; return d; }
XExtCodes* XAddExtension(Display* d){ _XExtension* e=calloc(1,sizeof(*e)); e->codes.extension=d->ext_number++; e->next=d->ext_procs; d->ext_procs=e; LOG("XAddExtension\n"); return &e->codes; }
XVisualInfo* XGetVisualInfo(Display* d,long mask,XVisualInfo* t,int* n){ LOG("XGetVisualInfo mask=%lx depth=%d class=%d id=%lx\n",mask,t?t->depth:-1,t?t->c_class:-1,t?t->visualid:0);
  if((mask&0x1)&&t->visualid!=g_vis.visualid){*n=0;return 0;} if((mask&0x2)&&t->screen!=0){*n=0;return 0;} if((mask&0x4)&&t->depth!=24){*n=0;return 0;} if((mask&0x8)&&t->c_class!=4){*n=0;return 0;}
  XVisualInfo* v=calloc(1,sizeof(*v)); v->visual=&g_vis; v->visualid=g_vis.visualid; v->screen=0; v->depth=24; v->c_class=4; v->red_mask=g_vis.red_mask; v->green_mask=g_vis.green_mask; v->blue_mask=g_vis.blue_mask; v->colormap_size=256; v->bits_per_rgb=8; *n=1; return v; }
int XFree(void* p){ free(p); return 1; }
static int img_destroy(XImage* i){ free(i); return 1; }
XImage* XCreateImage(Display* d,Visual* v,unsigned depth,int format,int offset,char* data,unsigned w,unsigned h,int pad,int bpl){ XImage* i=calloc(1,sizeof(*i)); i->width=w;i->height=h;i->format=format;i->data=data;i->bitmap_unit=32;i->bitmap_pad=pad;i->depth=depth;i->bits_per_pixel=32;i->bytes_per_line=bpl?bpl:(int)w*4;i->red_mask=0xff0000;i->green_mask=0xff00;i->blue_mask=0xff;i->f.destroy_image=img_destroy; LOG("XCreateImage %ux%u d=%u\n",w,h,depth); return i; }
Pixmap XCreatePixmap(Display* d,Drawable dr,unsigned w,unsigned h,unsigned depth){ Pix* p=&g_pix[g_npix++%4096]; p->id=g_next++; p->w=w;p->h=h;p->depth=depth; LOG("XCreatePixmap %ux%u -> %lx\n",w,h,p->id); return p->id; }
int XFreePixmap(Display* d,Pixmap p){ return 1; }
Status XGetGeometry(Display* d,Drawable dr,Window* root,int* x,int* y,unsigned* w,unsigned* h,unsigned* bw,unsigned* depth){ for(int i=0;i<g_npix&&i<4096;i++) if(g_pix[i].id==dr){ if(root)*root=g_scr.root; if(x)*x=0; if(y)*y=0; if(w)*w=g_pix[i].w; if(h)*h=g_pix[i].h; if(bw)*bw=0; if(depth)*depth=g_pix[i].depth; return 1;} LOG("XGetGeometry unknown %lx\n",dr); if(root)*root=g_scr.root; if(x)*x=0; if(y)*y=0; if(w)*w=16; if(h)*h=16; if(bw)*bw=0; if(depth)*depth=24; return 1; }
Colormap XCreateColormap(Display* d,Window w,Visual* v,int alloc){ return 0x30; }
GC XCreateGC(Display* d,Drawable dr,unsigned long m,void* v){ return &g_gc; }
int XFreeGC(Display* d,GC g){ return 1; }
int XSetFunction(Display* d,GC g,int f){ return 1; }
int XSetForeground(Display* d,GC g,unsigned long c){ return 1; }
int XFillRectangle(Display* d,Drawable dr,GC g,int x,int y,unsigned w,unsigned h){ return 1; }
int XPutImage(Display* d,Drawable dr,GC g,XImage* i,int sx,int sy,int dx,int dy,unsigned w,unsigned h){ return 1; }
XImage* XGetImage(Display* d,Drawable dr,int x,int y,unsigned w,unsigned h,unsigned long pm,int fmt){ return 0; }
int XFlush(Display* d){ return 1; } int XSync(Display* d,Bool b){ return 1; }
int (*XSynchronize(Display* d,Bool b))(Display*){ return 0; }
typedef int (*XErrorHandler)(Display*,void*); static XErrorHandler g_eh; XErrorHandler XSetErrorHandler(XErrorHandler h){ XErrorHandler o=g_eh; g_eh=h; return o; }
Bool XQueryExtension(Display* d,const char* name,int* a,int* b,int* c){ LOG("XQueryExtension %s\n",name); return 0; }
Status XGetWindowAttributes(Display* d,Window w,void* attr){ LOG("XGetWindowAttributes\n"); return 0; }
void* XQueryFont(Display* d,XID f){ return 0; } int XFreeFontInfo(char** n,void* i,int c){ return 1; } int XDrawString16(){ return 1; }
