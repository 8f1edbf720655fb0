#include "fakex.h"
Bool XShmAttach(Display* d,void* s){ return 0; } XImage* XShmCreateImage(){ return 0; } Bool XShmPutImage(){ return 0; }
