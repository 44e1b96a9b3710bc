// pg_sweep.h — host-side launchers of the kernels in pg_sweep.cu.
#pragma once
#include <cuda_runtime.h>
#include "pg_types.h"

int pg_query_grid(int device, int S, int smem_kb, int* grid_out, size_t* smem_out, int* stage_out, char* err, size_t errlen);
cudaError_t pg_launch_prepare(const PgState& st, int tune, cudaStream_t s);
cudaError_t pg_launch_persistent(const PgState& st, size_t smem, cudaStream_t s);
cudaError_t pg_launch_phase(const PgState& st, size_t smem, cudaStream_t s);
cudaError_t pg_launch_serial(const PgState& st, cudaStream_t s);
cudaError_t pg_launch_fill(double* p, long long n, double v, cudaStream_t s);
cudaError_t pg_launch_init_forest(const PgState& st, cudaStream_t s);
cudaError_t pg_launch_leaf_indices(const PgState& st, int t, unsigned* out, cudaStream_t s);
cudaError_t pg_launch_predict(const PgState& st, const float* x, long long n_new, long long ldx, double* out, cudaStream_t s);
