/* cuda.h -- empty companion of the CPU shim (the reference includes it, Deff2D.cuh:16,
 * but uses nothing from the driver API).  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_CUDA_SHIM_CUDA_H
#define ORACLE_CUDA_SHIM_CUDA_H
#endif
