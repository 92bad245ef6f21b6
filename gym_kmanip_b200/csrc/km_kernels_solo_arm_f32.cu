// Kernels of the solo_arm scene in float precision (one translation unit per instantiation so they compile in parallel).
#include "km_launch.cuh"
namespace km { KmVtable vtable_solo_arm_f32() { return Launch<SceneSoloArm, float>::vtable(); } }
