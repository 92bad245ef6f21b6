// Kernels of the torso scene in float precision (one translation unit per instantiation so they compile in parallel).
#include "km_launch.cuh"
namespace km { KmVtable vtable_torso_f32() { return Launch<SceneTorso, float>::vtable(); } }
