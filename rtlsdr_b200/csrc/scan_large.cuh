#pragma once
#include "scan_kernels.cuh"
