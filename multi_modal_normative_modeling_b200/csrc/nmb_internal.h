// Internal (non-ABI) declarations shared by the translation units of libnmb.
#pragma once
#include <string>
#include "nmb_common.cuh"
#include "nmb_tcp.h"

namespace nmb {

// Device-resident argument tables (arrays of per-segment pointers / sizes).
struct SegTable {
  const float* const* x;        // packed rows [n][ldx]
  const int* ldx;
  const float* const* xhat;     // [n][d]
  const uint8_t* const* mask;   // optional table (entries may be NULL)
  const int* n_rows;
  const int* d;
  float* const* stats_out;      // [2][d]
  const float* const* stats;    // optional
  float* const* dev_roi;        // optional
  float* const* z;              // optional
  float* const* dev_subj;       // optional
};

struct AucTable {
  const float* const* scores;
  const uint8_t* const* labels;
  const int* n_rows;
  const int* n_cols;
  double* const* out_auc;
  unsigned long long* const* out_u2;   // optional
};

struct PtrTable16 { const float* p[16]; };

void launch_pack_rows(const float* x, const float* c, long long n_rows, int d, int c_dim, int ldx, float* out,
                      cudaStream_t st);
void launch_stats(const SegTable& t, int n_seg, int max_d, cudaStream_t st);
void launch_deviation(const SegTable& t, int n_seg, int max_rows, cudaStream_t st);
void launch_auc(const AucTable& t, int n_seg, int max_cols, int max_rows, cudaStream_t st);
struct LatentTable {
  const float* const* mu_train; const int* n_train; const float* const* mu; const float* const* logvar;
  const int* n_rows; const int* latent; float* const* out_z; float* const* out_dev;
};
void launch_latent_deviation(const LatentTable& t, int n_seg, int max_rows, cudaStream_t st);
int sm_count();
// nmb_prologue.cu
int prologue_max_rows();
cudaError_t launch_robust_fit(const double* x, long long ld, int d, const int* idx, int n, double* center, double* scale,
                              cudaStream_t st);
cudaError_t launch_rank_bins(const double* v, const int* idx, int n, const double* edges, int q, int* bins, cudaStream_t st);
cudaError_t launch_pack_scaled(const double* x, long long ld, int d, const int* idx, long long n, const double* center,
                               const double* scale, const int* age_bin, int n_age, const int* sex_bin, int n_sex, int ldx,
                               float* out, cudaStream_t st);
void launch_mean_rows(const PtrTable16& src, int k, long long n, float* out, cudaStream_t st);
void launch_member_records(const float* stats, const long long* o_stats, const double* auc_roi, const long long* o_auc,
                           const double* auc_subj, const float* subj, const long long* o_subj, const int* seg_d, const int* n_test,
                           int n_seg, int d_max, int n_test_max, double* out, cudaStream_t st);
void launch_adam(float* p, const float* g, float* m, float* v, long long n, float step_size, float bc2_sqrt,
                 float b1, float b2, float eps, cudaStream_t st);
void launch_philox(unsigned long long seed, unsigned long long step, uint32_t stream_id, long long n, float* out,
                   cudaStream_t st);

// nmb_train.cu
struct TrainLaunch {
  MemberDev* members; const ArchDesc* archs; int n_members;
  long long n_steps; const float* eps_override; float* loss_out; unsigned flags;
  int n_is_epochs;          // 1: member i runs n_steps * steps_per_epoch_i steps (nmb_ensemble_train_epochs)
  long long stride_steps;   // rows per member of eps_override / loss_out
  float* scratch; long long slot_floats; int n_slots; int* work_counter;
  const int* order;   // members sorted by decreasing cost (longest-processing-time-first dealing)
};
// minibatch steps member `mb` takes in this launch
__host__ __device__ inline long long member_steps(const TrainLaunch& t, const MemberDev& mb) {
  return t.n_is_epochs ? t.n_steps * (long long)((mb.n_rows + mb.batch - 1) / mb.batch) : t.n_steps;
}
cudaError_t launch_train(const TrainLaunch& t, cudaStream_t st);
cudaError_t launch_debug_tc_gemm(const float* A, int lda, int a_kmajor, const float* B, int ldb, int b_kmajor,
                                 float* C, int ldc, int M, int N, int K, cudaStream_t st);

struct ReconItem { int member; int row0; int rows; };
struct ReconLaunch {
  MemberDev* members; const ArchDesc* archs;
  const ReconItem* items; int n_items;
  const float* const* xc; const float* const* eps; float* const* xhat; float* const* mu; float* const* logvar;
  int mode; float* scratch; long long slot_floats; int n_slots;
  int fp32;   // FP32 FFMA engine instead of tcgen05
  float* const* head_out;   // optional per member: head predictions [n_rows] (members with a head)
};
cudaError_t launch_recon(const ReconLaunch& t, cudaStream_t st);
cudaError_t configure_kernels();

// nmb_train_tcp.cu: pipelined tensor-core path
cudaError_t configure_tcp();
cudaError_t set_tcp_trace(unsigned long long* buf, int step);
cudaError_t launch_xprep(const void* items_dev, int n_items, int max_blocks, cudaStream_t st);
cudaError_t launch_recon_tcp(const TrainLaunch& t, const tcp::ProgramDev* progs, const tcp::ProgramDev* train_progs,
                             const tcp::MemberTc* mtc, unsigned char* stash, long long stash_bytes,
                             const tcp::MStep* msteps, const int* ms_off, const int* ms_cnt, int n_archs,
                             const tcp::EpiP* epis_p, const int* ep_off, const int* ep_cnt, const unsigned char* ep_first,
                             int max_mlayers, int recon_mode, const tcp::ReconTc* rtc, const tcp::ReconWork* rwork, int n_rwork,
                             int n_sm, cudaStream_t st);
cudaError_t launch_train_tcp(const TrainLaunch& t, const tcp::ProgramDev* progs, const tcp::MemberTc* mtc,
                             unsigned char* stash, long long stash_bytes, float* master, long long master_floats,
                             const tcp::MStep* msteps, const int* ms_off, const int* ms_cnt, int n_archs,
                             const tcp::EpiP* epis_p, const int* ep_off, const int* ep_cnt, const unsigned char* ep_first,
                             int max_mlayers, int n_sm, bool gather_in, bool scatter_out, cudaStream_t st);
cudaError_t launch_tcp_scatter(MemberDev* members, int n_members, const tcp::ProgramDev* progs, const tcp::MemberTc* mtc,
                               float* master, long long master_floats, int max_mlayers, cudaStream_t st);

// nmb_csv.cu: host-side table writer (f2)
int csv_write(const char* path, const char* header, const char* const* prefix, const void* body, int is_f64, long long n_rows,
              int n_cols, long long ld, int threads, std::string& err);

}  // namespace nmb
