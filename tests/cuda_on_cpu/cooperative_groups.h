// TEST INFRASTRUCTURE ONLY -- cooperative_groups (clusters) for the "CUDA on CPU" shim: see common.cuh
#pragma once
#include "common.cuh"
