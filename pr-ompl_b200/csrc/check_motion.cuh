// Internal C++ interface of the K2 kernels; see check_motion.cu.
#pragma once
#include "world.cuh"
namespace prgpu {
struct CmScratch { // owed-state counts, their prefix sums and the scan's temporary storage
  uint32_t *owed = nullptr, *off = nullptr;
  void *tmp = nullptr;
  size_t cap = 0, tmp_bytes = 0;
  uint2 *gaps = nullptr;         // per edge 8 x (first owed state, count): what phase A left untested
  uint2 *queue = nullptr;        // samples the clearance field could not certify: (edge * 8 + slot, sphere mask)
  uint32_t *queue_n = nullptr;
  uint32_t *item_edge = nullptr; // owed state -> edge, followed by owed state -> tested-state index
  size_t item_cap = 0;
  // pair-dealt pipeline: one item per (state, uncertified link sphere)
  uint32_t *pair_ref = nullptr;
  void *pair_centre = nullptr;   // float4: centre + sphere slot
  uint32_t *pair_n = nullptr;    // item counters of phase A / phase B
  size_t pair_cap = 0;
};
int check_motion_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const double *from_dev, const double *to_dev, uint64_t n,
                        double resolution, int mode, int check_first, uint8_t *valid_dev, int sm_count,
                        cudaStream_t stream, KernelTimer *timer = nullptr, CmScratch *scratch = nullptr);
// clearance field (world.cuh): df[cell] = min over obstacles of the fp64 distance from the cell's box, capped at
// `cap`, minus `slack`, rounded down to fp32
int dfield_build_launch(const double *os_d, int n_os, const double *ob_d, int n_ob, const int *n3, const double *org,
                        double cell, double slack, double cap, uint32_t *df_dev, cudaStream_t stream);
int is_valid_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const double *q_dev, uint64_t n, uint8_t *valid_dev,
                    cudaStream_t stream);
int diag_fk_centre_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const double *q_dev, uint64_t n, double *err_dev,
                          cudaStream_t stream);
int fk_launch(const DeviceWorld &w, const DeviceWorld *w_dev, const double *q_dev, uint64_t n, double *pose12_dev, cudaStream_t stream);
} // namespace prgpu
