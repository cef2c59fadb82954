// cabm_kernels.h — host-callable launchers of the kernels in cabm_kernels.cu. Each returns the number of
// kernel launches it issued (for cabm_launch_count).
#pragma once
#include "cabm_device.cuh"

namespace cabm {

void upload_const_tables(const ConstTables &t, cudaStream_t s);
int launch_initialize(const Dev &S, cudaStream_t s);
int launch_build_groups(const Dev &S, cudaStream_t s);
int launch_insert_infected(const Dev &S, int which, int health, int num, int ag, int call_id, int day, cudaStream_t s);
int launch_first_column(const Dev &S, cudaStream_t s);
int launch_dyntrans_exact(const Dev &S, int day, cudaStream_t s);
int launch_dyntrans_native(const Dev &S, int day, cudaStream_t s);
int launch_ct_identify(const Dev &S, int day, bool scan, cudaStream_t s);
int launch_time_update_scan(const Dev &S, int day, bool any_ct, cudaStream_t s);
int launch_time_update_events(const Dev &S, int day, cudaStream_t s);
int launch_curves_finalize(const Dev &S, cudaStream_t s);
int launch_count_infectors(const Dev &S, cudaStream_t s);
int launch_reduce_hmatrix(const int16_t *hm, const int *ags, int M, int N, int T, int *inc, int *prev, cudaStream_t s);
int launch_widen(const uint8_t *src, int16_t *dst, size_t n, cudaStream_t s);
int launch_sim_fused(const Dev &S, unsigned *work_counter, int n_sm, bool any_ct, bool global_state, cudaStream_t s);
size_t fused_smem_bytes_host(int Np, bool global_state);
int launch_narrow(const int *a, const int *b, int16_t *out, size_t n, cudaStream_t s);
int launch_mean_curves(const int *inc, const int *prev, double *minc, double *mprev, int R, size_t per_rep, cudaStream_t s);
int launch_peek(const Dev &S, int field, int r, int day, int *out, cudaStream_t s);
int launch_poke(const Dev &S, int field, int r, int i, int day, int value, cudaStream_t s);

}  // namespace cabm
