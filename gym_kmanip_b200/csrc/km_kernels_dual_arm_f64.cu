// Kernels of the dual_arm scene in double precision (one translation unit per instantiation so they compile in parallel).
#include "km_launch.cuh"
namespace km { KmVtable vtable_dual_arm_f64() { return Launch<SceneDualArm, double>::vtable(); } }
