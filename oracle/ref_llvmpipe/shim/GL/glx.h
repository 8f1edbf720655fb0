#pragma once
/* minimal stand-in for <GL/glx.h>: gl3w.c only needs glXGetProcAddress */
#ifdef __cplusplus
extern "C" {
#endif
typedef void (*__GLXextFuncPtr)(void);
__GLXextFuncPtr glXGetProcAddress(const unsigned char*);
#ifdef __cplusplus
}
#endif
