// Kernels of the solo_arm scene in double precision (one translation unit per instantiation so they compile in parallel).
#include "km_launch.cuh"
namespace km { KmVtable vtable_solo_arm_f64() { return Launch<SceneSoloArm, double>::vtable(); } }
