// lf_mesh_bvh_hook.h — force-included (after ref_prelude.h) in front of the reference's UNCHANGED Scene.cpp, and of nothing else.
//
// Scene.cpp:31 is the reference's one `new Mesh`, so this translation unit is the only one that instantiates the inline constructor of
// Mesh.h:16-20, whose body is `bvh = new RadeonRays::SplitBvh(2.0f, 64, 0, 0.001f, 0);`.  With split_bvh.h already included (its include guard
// keeps Mesh.h from including it again) the macro below makes that one expression construct the subclass of DeviceBvh.h, which behaves
// exactly like SplitBvh unless the device build is switched on.  Scene.cpp names SplitBvh nowhere else.  No reference source is edited.
#pragma once
#include "split_bvh.h"
#include "DeviceBvh.h"
#define SplitBvh LfDeviceSplitBvh
