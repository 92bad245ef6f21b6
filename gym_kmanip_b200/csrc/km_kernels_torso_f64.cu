// Kernels of the torso scene in double precision (one translation unit per instantiation so they compile in parallel).
#include "km_launch.cuh"
namespace km { KmVtable vtable_torso_f64() { return Launch<SceneTorso, double>::vtable(); } }
