// task_dev.cuh -- kernel-parameter structs shared by the dynamics and the env kernels (and by every translation unit
// that instantiates them).
#pragma once
#include "cuda_compat.cuh"
#include "../../include/b200_step.h"

struct TaskDev {
    float f[TF_COUNT];
    int i[TI_COUNT];
};

struct ModelDev {          // device pointers to the packed robot model (see robot_model.py)
    const float *body;     // [nb][20]
    const float *link_off; // [nlinks][3]
    const float *sph;      // [nspheres][4]
    const int *link_body;  // [nlinks]
    const int *sph_body;   // [nspheres]
    const int *sph_link;   // [nspheres]
};

struct TerrainDev {
    const int16_t *hf;     // [rows][cols] or nullptr (plane)
    const float *origins;  // [levels][types][3]
    int rows, cols, levels, types;
};
