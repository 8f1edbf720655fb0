// host_globals.cpp — the two definitions the reference's host sources expect from their application:
// `LavaFrameState GlobalState` (declared extern in Scene.cpp:11, Mesh.cpp:13, Loader.cpp:37; defined
// by Main.cpp in the reference) and the single stb_image implementation (Export.h:2 in the reference).
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#include "GlobalState.h"

LavaFrameState GlobalState;
