// Pipelined tensor-core training kernel (see nmb_tcp.h for the design and the storage format).
//
// Reference behaviour replaced (file:line in soz223/multi_modal_normative_modeling):
//   hot loop                     multimodal_kfold_train_cvae_supervised.py:177-199
//   Encoder/Decoder.forward      cVAE.py:161-172, 197-206
//   combine_latent + experts     cVAE.py:1144-1164, 986-1083
//   reparameterise, KL, LL       cVAE.py:1130-1133, 1138-1139, 14-15   (-MSE: ..._nmmlp.py:124-127)
//   loss_function_multimodal     cVAE.py:1187-1196
//   optimizer1 = Adam(...)       cVAE.py:1111-1116 (torch defaults)
// One persistent CTA per SM; a member's minibatch steps all run inside one launch.
#include <cstdlib>
#include <cstring>

#include "nmb_tc_gemm.cuh"
#include "nmb_internal.h"
#include "nmb_tcp.h"
#include "nmb_fusion.cuh"

namespace nmb {
namespace tcp {

// Optional timeline trace (test hook): globaltimer stamps of one minibatch step of CTA 0.
__device__ unsigned long long* g_trace = nullptr;
__device__ int g_trace_step = 2;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(cond, idx) do { if (cond) g_trace[idx] = gtime(); } while (0)
// Fine-grained stamps inside the items (dev builds only: -DNMB_TCP_FINE_TRACE; SM clock cycles of one thread).
#ifdef NMB_TCP_FINE_TRACE
#define TR3() do { if (c.tr_ptr) *c.tr_ptr++ = (unsigned long long)clock64(); } while (0)   // per-thread cursor in EpiCtx
#else
#define TR3() do { } while (0)
#endif

// Rows walked by the current work item: the member's training rows (minibatch `batch`) or, in a reconstruction launch,
// the caller's rows in tiles of 256.  Lives in shared memory (Ctrl), read per step: no role keeps it in registers.
struct RowSet { int n_rows, batch, n_half; };

struct Ctrl {
  uint64_t full[kSlots], empty[kSlots], accbar[6];      // accbar: acc[0], acc[1], wacc[0], wacc[1], upper halves of acc[0], acc[1]
  uint32_t tmem;
  volatile uint32_t epi_done[kGroups];   // per epilogue group: items finished (1 + step * n_epis + index + 1)
  int member;
  int chunk_i0, chunk_n;                 // launch-relative first step and step count of the current work item
  long long chunk_s0;                    // member-global index of its first step
  RowSet rs;                             // rows of the current work item
  int rt_idx;                            // reconstruction launch: index of the work item's row set (ReconTc)
  float red[40];
};

// Item counters: release store by the publishing thread, acquire load by the waiting role (CTA scope) -- no
// separate MEMBAR on either side.
__device__ __forceinline__ uint32_t ld_acquire(const volatile uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(tc::smem_u32((const void*)p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(volatile uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(tc::smem_u32((const void*)p)), "r"(v) : "memory");
}
// Bounded spin with back-off (a protocol bug must trap, not hang the GPU; the spinning single-thread
// warps must not steal issue slots from the epilogue warps of their scheduler).
#ifndef NMB_POLL_NS
#define NMB_POLL_NS 100
#endif
constexpr unsigned kPollSleepNs = NMB_POLL_NS;
// Watchdog of every bounded spin, in SM clock cycles (0 = never trap).  Set per process from NMB_TCP_SPIN_BUDGET
// (configure_tcp): tools that stretch time by orders of magnitude -- compute-sanitizer, ncu replay, a debugger -- must
// be able to switch it off, because a trap leaves a sticky context error.  Read only on the slow path.
__device__ long long g_spin_budget = 8000000000LL;
__device__ __forceinline__ void wait_epi(const volatile uint32_t* p, uint32_t need) {
  if (ld_acquire(p) < need) {
    const long long t0 = clock64();
    while (ld_acquire(p) < need) {
      __nanosleep(kPollSleepNs);
      if (g_spin_budget > 0 && clock64() - t0 > g_spin_budget) __trap();
    }
  }
}
__device__ __forceinline__ void wait_all(const volatile uint32_t* p, uint32_t need) {
#pragma unroll
  for (int g = 0; g < kGroups; ++g) wait_epi(p + g, need);
}
__device__ __forceinline__ void bar_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc::smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

struct StepVars {           // per minibatch step, identical in every role
  int rows, rows_h[2], row0, pos;
  uint32_t base;            // epi_done value when every item of the previous step has finished
};
__device__ __forceinline__ StepVars step_vars(const RowSet& rs, long long s, long long i, int n_epis) {
  StepVars v;
  const int spe = (rs.n_rows + rs.batch - 1) / rs.batch;
  v.pos = (int)(s % spe);
  v.row0 = v.pos * rs.batch;
  v.rows = min(rs.batch, rs.n_rows - v.row0);
  v.rows_h[0] = min(v.rows, 128);
  v.rows_h[1] = v.rows - v.rows_h[0];
  v.base = 1u + (uint32_t)i * (uint32_t)n_epis;
  return v;
}

// ------------------------------------------------------------------------------------------------
// producer: TMA bulk copies of operand tiles, in program order, into the ring
// The fields of a Step the producer needs, fetched one step ahead (the table lives in global memory: an L2 round
// trip per step would otherwise sit on the single producer thread's critical path).
struct PStep {
  long long b_off, a_off;
  unsigned b_bytes, a_bytes;
  int dep;
  unsigned char half, b_space, a_space, x_mod, dep_grp;
};
__device__ __forceinline__ PStep load_pstep(const Step* steps, int k) {
  const Step& st = steps[k];
  PStep p;
  p.b_off = st.b_off; p.a_off = st.a_off; p.b_bytes = st.b_bytes; p.a_bytes = st.a_bytes; p.dep = st.dep;
  p.half = st.half; p.b_space = st.b_space; p.a_space = st.a_space; p.x_mod = st.x_mod; p.dep_grp = st.dep_grp;
  return p;
}

__device__ void producer_role(const TrainLaunch& t, const ProgramDev& pg, const RowSet& rs, const MemberTc& mt,
                              const unsigned char* const* xplanes, bool recon,
                              const unsigned char* stash, unsigned char* smem, Ctrl* ctl, uint32_t& seq) {
  unsigned char* ring = smem + kSmemRing;
  const Step* __restrict__ steps = pg.steps;
  const int n_steps = pg.n_steps;
  const long long s0 = ctl->chunk_s0;
  const int i0 = ctl->chunk_i0, n_chunk = ctl->chunk_n;
  fence_async_all();          // weight planes may have been written by another SM (previous work item of this member)
  PStep nxt = load_pstep(steps, 0);
  for (long long i = 0; i < n_chunk; ++i) {
    const StepVars sv = step_vars(rs, s0 + i, i, pg.n_epis);
    const bool tr = g_trace && blockIdx.x == 0 && i0 + i == g_trace_step;
    const int tb = 3 * pg.n_epis + 3 * pg.n_steps;
    for (int k = 0; k < n_steps; ++k) {
      const PStep st = nxt;
      nxt = load_pstep(steps, k + 1 < n_steps ? k + 1 : 0);
      if (sv.rows_h[st.half] == 0) continue;
      if (st.dep) {              // data written by an epilogue item of this step (stash blocks behind their fence)
        if (st.dep_grp < 2) wait_epi(&ctl->epi_done[st.dep_grp], sv.base + (uint32_t)st.dep);
        else wait_all(ctl->epi_done, sv.base + (uint32_t)st.dep);
      }
      TRACE(tr, tb + 2 * k);
      for (int which = 0; which < 2; ++which) {
        const int space = which == 0 ? st.a_space : st.b_space;
        const uint32_t bytes = which == 0 ? st.a_bytes : st.b_bytes;
        if (bytes == 0) continue;          // A resident in ACT[half] / B = the tile the previous step kept
        // weight planes: every Adam item of the previous step (dataset tiles do not wait: the first x tile of a step
        // is in flight while the previous step ends)
        // (a reconstruction launch never rewrites the planes: its weight tiles run ahead of the previous tile's items)
        if (space == SP_W && !recon) wait_all(ctl->epi_done, sv.base);
        const long long off = which == 0 ? st.a_off : st.b_off;
        const unsigned char* src;
        if (space == SP_W) src = mt.wplanes + off;
        else if (space == SP_STASH) src = stash + off;
        else src = xplanes[st.x_mod] + ((long long)(sv.pos * rs.n_half + st.half) * pg.lay.x_cg[st.x_mod]) * 4096 + off;
        const uint32_t slot = seq % kSlots, use = seq / kSlots;
        tc::mbar_wait(&ctl->empty[slot], (use & 1u) ^ 1u);
        expect_tx(&ctl->full[slot], bytes);
        bulk_load(ring + slot * kSlotBytes, src, bytes, &ctl->full[slot]);
        ++seq;
      }
      TRACE(tr, tb + 2 * k + 1);
    }
  }
}

struct LaunchP {
  TrainLaunch t;
  const ProgramDev* progs;
  const MemberTc* mtc;
  unsigned char* stash;
  long long stash_bytes;
  float* master;            // per slot: 3 x master_floats (p, m, v)
  long long master_floats;
  int ms_off[kMaxParamArchs], ms_cnt[kMaxParamArchs];   // slice of msteps per architecture
  int n_chunks;             // a member's steps of this launch are dealt as n_chunks work items (consecutive step ranges)
  // forward-only (reconstruction) launch: recon_mode 1 = decode the mean, 2 = sampled z; work items = tile ranges
  int recon_mode; int n_rwork;
  const ReconTc* rtc; const ReconWork* rwork;
  int ep_off[kMaxParamArchs], ep_cnt[kMaxParamArchs];   // slice of epis_p per architecture; cnt 0 = global-memory table
  unsigned char ep_first[kMaxParamArchs][4];            // first own item of epilogue group g (see EpiP::next_own)
  int jump;                                             // 0: dev knob NMB_TCP_JUMP=0, every group walks the whole item list
  MStep msteps[kMaxParamSteps];
  EpiP epis_p[kMaxParamEpis];
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------------------------
// MMA issuer: 3 BF16 passes per K = 16 step.  The WHOLE warp walks the step table (kernel parameters ->
// uniform loads); one elected lane issues, so descriptors never leave the uniform datapath.
__device__ void mma_role(const LaunchP& L, int ai, const RowSet& rs, unsigned char* smem, Ctrl* ctl, uint32_t tmem,
                         uint32_t& seq, int n_epis) {
  const uint32_t ring = tc::smem_u32(smem + kSmemRing);
  const uint32_t act0 = tc::smem_u32(smem);
  const long long s0 = ctl->chunk_s0;
  const int i0 = ctl->chunk_i0, n_chunk = ctl->chunk_n;
  const int k0 = L.ms_off[ai], k1 = k0 + L.ms_cnt[ai];
  uint32_t held_slot = 0;
  for (long long i = 0; i < n_chunk; ++i) {
    const StepVars sv = step_vars(rs, s0 + i, i, n_epis);
    const bool half1 = __any_sync(0xffffffffu, sv.rows_h[1] > 0);
    wait_all(ctl->epi_done, sv.base);
    const bool tr = g_trace && blockIdx.x == 0 && i0 + i == g_trace_step;
    const int tb = 3 * n_epis;
    for (int k = k0; k < k1; ++k) {
      const MStep& st = L.msteps[k];
      if (st.half && !half1) continue;
      if (st.mma_dep) wait_epi(&ctl->epi_done[st.half], sv.base + (uint32_t)st.mma_dep);
      if (st.mma_dep_joint > 0) wait_epi(&ctl->epi_done[2], sv.base + (uint32_t)st.mma_dep_joint);
      else if (st.mma_dep_joint < 0) wait_all(ctl->epi_done, sv.base + (uint32_t)(-st.mma_dep_joint));
      if (tr && (threadIdx.x & 31) == 0) g_trace[tb + 3 * (k - k0)] = gtime();
      uint32_t a_base, slot_a = kSlots, slot_b0 = kSlots;
      if (st.a_tile & 4) {         // first K half of B (the step's first ring tile); A stays resident in ACT[half]
        slot_b0 = seq % kSlots;
        tc::mbar_wait(&ctl->full[slot_b0], (seq / kSlots) & 1u);
        ++seq;
        a_base = act0 + st.half * kActBytes + st.a_start;
      } else if (st.a_tile & 1) {
        slot_a = seq % kSlots;
        tc::mbar_wait(&ctl->full[slot_a], (seq / kSlots) & 1u);
        a_base = ring + slot_a * kSlotBytes;
        ++seq;
      } else {
        a_base = act0 + st.half * kActBytes + st.a_start;
      }
      uint32_t slot_b;
      if (st.b_mn & 2) {           // B = the tile the previous step kept (already landed): released after these MMAs
        slot_b = held_slot;
      } else {
        slot_b = seq % kSlots;
        tc::mbar_wait(&ctl->full[slot_b], (seq / kSlots) & 1u);
        ++seq;
      }
      const uint32_t b_base = ring + slot_b * kSlotBytes;
      const bool keep_a = (st.a_tile & 2) != 0;
      if (keep_a) held_slot = slot_a;
      tc::fence_after();
      const uint32_t idesc = tc::make_idesc(st.n, st.a_mn, st.b_mn & 1);
      const uint32_t d = tmem + st.tmem_col;
      const uint64_t hi_a = ((uint64_t)st.a_sbo << 32) | (1ull << 46) | ((uint64_t)st.a_lbo << 16);
      const uint64_t hi_b = ((uint64_t)st.b_sbo << 32) | (1ull << 46) | ((uint64_t)st.b_lbo << 16);
      uint64_t da = hi_a | ((a_base & 0x3FFFFu) >> 4), db = hi_b | ((b_base & 0x3FFFFu) >> 4);
      if (elect_one()) {
        if (tr) g_trace[tb + 3 * (k - k0) + 1] = gtime();
        uint32_t acc = st.first ? 0u : 1u;
        if (slot_b0 != kSlots) {       // two-part step: k_first K steps on the first B tile, the rest on the second
          uint64_t db0 = hi_b | (((ring + slot_b0 * kSlotBytes) & 0x3FFFFu) >> 4);
          const int k1 = st.k_first;
          for (int ks = 0; ks < k1; ++ks) {
            tc::mma_bf16(d, da, db0, idesc, acc);
            tc::mma_bf16(d, da + st.a_lo, db0, idesc, 1u);
            tc::mma_bf16(d, da, db0 + st.b_lo, idesc, 1u);
            acc = 1u;
            da += st.a_kadv; db0 += st.b_kadv;
          }
          tc::mma_commit(&ctl->empty[slot_b0]);
          uint32_t d2 = d, idesc2 = idesc;
          if (st.p2_new) {             // a new N chunk: A from the start again, its own accumulator columns and width
            da = hi_a | ((a_base & 0x3FFFFu) >> 4);
            d2 = tmem + st.tmem2; idesc2 = tc::make_idesc(st.n2, st.a_mn, st.b_mn & 1);
            acc = st.first ? 0u : 1u;
          }
          for (int ks = k1; ks < st.ksteps; ++ks) {
            tc::mma_bf16(d2, da, db, idesc2, acc);
            tc::mma_bf16(d2, da + st.a_lo, db, idesc2, 1u);
            tc::mma_bf16(d2, da, db + st.b_lo, idesc2, 1u);
            acc = 1u;
            da += st.a_kadv; db += st.b_kadv;
          }
        } else
        for (int ks = 0; ks < st.ksteps; ++ks) {
          tc::mma_bf16(d, da, db, idesc, acc);
          tc::mma_bf16(d, da + st.a_lo, db, idesc, 1u);
          tc::mma_bf16(d, da, db + st.b_lo, idesc, 1u);
          acc = 1u;
          da += st.a_kadv; db += st.b_kadv;
        }
        tc::mma_commit(&ctl->empty[slot_b]);
        if (slot_a != kSlots && !keep_a) tc::mma_commit(&ctl->empty[slot_a]);
        if (st.commit == 1 || (st.commit == 2 && !half1)) tc::mma_commit(&ctl->accbar[st.commit_buf]);
        if (st.commit2) tc::mma_commit(&ctl->accbar[st.half]);
        if (tr) g_trace[tb + 3 * (k - k0) + 2] = gtime();
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// epilogue helpers
struct EpiCtx {
  const ArchDesc* a; const ProgramDev* pg; MemberDev* mb; const MemberTc* mt;
  unsigned char* smem; unsigned char* stash; float* scratch; Ctrl* ctl;
  float* mst_p; float* mst_m; float* mst_v;     // lane-major Adam master state of the resident member
  uint32_t tmem;
  int warp, lane, row;      // row = TMEM lane of this thread: 32 * (warp % 4) + lane
  int grp;                  // epilogue group = minibatch half this warp serves
  int cpart, parts;         // column partition of the current item (within the group, or across both for joint items)
  int tid, nthr;            // thread index / count of the current item's worker set
  int bar_id, bar_nthr;     // named barrier of that worker set
  unsigned flags;
  int rows, rows_h0, rows_h1, row0, pos;
  long long step;           // global 0-based minibatch step of this member (Philox counter, Adam t - 1)
  float step_size, inv_bc2, b1, b2, aeps;
  float kl_acc, ll_acc;
  float* dw_acc;            // [NMB_MAX_MOD] gPoE alpha-gradient partials (local array of the role)
  int recon;                // 0 = training step; 1 = reconstruction, decode the mean; 2 = reconstruction, sampled z
  const ReconTc* rt;        // rows and outputs of a reconstruction launch
  // row source of the current work item: the member's training rows, or the rows of a reconstruction launch (looked up
  // on demand: in the training instantiation `recon` is the constant 0 and nothing extra stays live in registers)
  template <bool R> __device__ __forceinline__ const float* xc_rows(int m) const { return R ? rt->xc[m] : mb->xc[m]; }
  template <bool R> __device__ __forceinline__ const unsigned char* cplanes0() const { return R ? rt->cplanes[0] : mt->cplanes[0]; }
  template <bool R> __device__ __forceinline__ int n_half() const { return R ? 2 : mt->n_half; }
  // reconstruction mode as a compile-time 0 in the training instantiation
  template <bool R> __device__ __forceinline__ int rmode() const { return R ? recon : 0; }
#ifdef NMB_TCP_FINE_TRACE
  unsigned long long* tr_ptr;
#endif
  __device__ __forceinline__ int rows_of(int h) const { return h ? rows_h1 : rows_h0; }
};

// 8 values -> hi / lo planes at (group g, row) of a 128-row block
__device__ __forceinline__ void put_planes(unsigned char* blk, int g, int row, const float (&x)[8]) {
  uint4 h, l;
  tc::split8(x, h, l);
  unsigned char* p = blk + (long long)g * 4096 + row * 16;
  *reinterpret_cast<uint4*>(p) = h;
  *reinterpret_cast<uint4*>(p + 2048) = l;
}

// Adam master state is a pure stream (read once, written once per step, re-read a whole step later): keep it from
// displacing the weight planes and the backward stash in L2.
// (.cg: the previous work item of the member may have run on another SM -- never trust this SM's L1 for it)
__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_stream4(float* p, const float4& v) { __stcs(reinterpret_cast<float4*>(p), v); }

// torch.optim.Adam element update; sqrt / reciprocal on the SFU (2 ulp, far inside the parity budget)
__device__ __forceinline__ float adam_update(const EpiCtx& c, float& m1, float& v1, float p0, float g) {
  m1 = c.b1 * m1 + (1.f - c.b1) * g;
  v1 = c.b2 * v1 + (1.f - c.b2) * g * g;
  float sq, rc;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(v1));
  const float denom = fmaf(sq, c.inv_bc2, c.aeps);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(denom));
  return fmaf(-c.step_size * m1, rc, p0);
}
__device__ __forceinline__ void adam_scalar(const EpiCtx& c, long long idx, float g) {
  MemberDev& mb = *c.mb;
  if (c.flags & NMB_TRAIN_WRITE_GRADS) mb.grads[idx] = g;
  if (!(c.flags & NMB_TRAIN_NO_ADAM)) {
    float m1 = __ldcg(mb.adam_m + idx), v1 = __ldcg(mb.adam_v + idx);
    const float p1 = adam_update(c, m1, v1, __ldcg(mb.params + idx), g);
    mb.adam_m[idx] = m1; mb.adam_v[idx] = v1; mb.params[idx] = p1;
  }
}

__device__ __forceinline__ uint32_t taddr(const EpiCtx& c, int col) {
  return c.tmem + ((uint32_t)((c.warp & 3) << 5) << 16) + (uint32_t)col;
}

// sum over ALL epilogue threads (joint items only: both groups call it)
__device__ __forceinline__ float block_sum_epi(EpiCtx& c, float v) {
  v = warp_sum(v);
  bar_n(4, kEpiWarps * 32);
  if (c.lane == 0) c.ctl->red[c.warp] = v;
  bar_n(4, kEpiWarps * 32);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kEpiWarps; ++i) s += c.ctl->red[i];
  return s;
}

// fp32 parameters -> BF16 hi/lo planes (member start)
// fp32 parameters -> BF16 hi/lo planes of the layer whose augmented matrix starts at float offset p_off (< 0: all)
__device__ void build_weight_planes(const ProgramDev& pg, const MemberDev& mb, const MemberTc& mt, int tid, int nthr,
                                    long long p_off) {
  const float* __restrict__ P = mb.params;
  for (int b = 0; b < pg.n_wblocks; ++b) {
    const WBlock wb = pg.wblocks[b];
    if (p_off >= 0 && wb.p_off != p_off) continue;
    if (wb.transposed) {            // rows = input index i, groups of 8 output rows
      unsigned char* dstT = mt.wplanes + wb.wp_off;
      for (int u = tid; u < wb.R * wb.cg; u += nthr) {
        const int i = u % wb.R, og = u / wb.R;
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          x[j] = (8 * og + j < wb.rows_valid && i < wb.cols_valid) ? P[wb.p_off + (long long)(8 * og + j) * wb.p_ld + i] : 0.f;
        uint4 h, l;
        tc::split8(x, h, l);
        unsigned char* p = dstT + (long long)og * 32 * wb.R + i * 16;
        *reinterpret_cast<uint4*>(p) = h;
        *reinterpret_cast<uint4*>(p + 16 * wb.R) = l;
      }
      continue;
    }
    const int units = wb.R * wb.cg;
    unsigned char* dst = mt.wplanes + wb.wp_off;
    for (int u = tid; u < units; u += nthr) {
      const int r = u % wb.R, g = u / wb.R;
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = 0.f;
      if (r < wb.rows_valid) {
        const float* src = P + wb.p_off + (long long)(wb.row0 + r) * wb.p_ld + 8 * g;
        const int nv = wb.cols_valid - 8 * g;
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nv) x[j] = src[j];
      }
      uint4 h, l;
      tc::split8(x, h, l);
      unsigned char* p = dst + (long long)g * 32 * wb.R + r * 16;
      *reinterpret_cast<uint4*>(p) = h;
      *reinterpret_cast<uint4*>(p + 16 * wb.R) = l;
    }
  }
}

template <bool RECON>
__device__ __forceinline__ void epi_hidden(EpiCtx& c, const Epi& e) {
  const int h = e.half;
  const int rows = c.rows_of(h);
  const bool vr = c.row < rows;
  unsigned char* act = c.smem + h * kActBytes + c.row * 16;
  // stash block: one 128-row block (group chunk 4 KB = hi 2 KB + lo 2 KB) or, src_cg > 0, two 64-row sub-blocks of
  // src_cg groups (group chunk 2 KB = hi 1 KB + lo 1 KB): the weight gradient that reads it is split over K
  const int b64 = e.src_cg;
  constexpr bool keep = !RECON;                   // forward-only program: nothing is stashed for a backward pass
  unsigned char* st = c.stash + (keep ? e.stash_off : 0LL) + (b64 ? (long long)(c.row >> 6) * b64 * 2048 + (c.row & 63) * 16 : (long long)c.row * 16);
  const int st_g = b64 ? 2048 : 4096, st_lo = b64 ? 1024 : 2048;
  const float slope = c.a->non_linear ? kSlope : 1.f;        // leaky-relu(x) = max(x, slope * x)
  const int n_cols = e.n_cols, n_mma = e.n_mma, n_valid = e.n_valid, tcol = e.tmem_col;
  for (int ch = c.cpart; ch * 16 < n_cols; ch += c.parts) {
    const int col = ch * 16;
    float v[16];
    if (col < n_mma) tc::tmem_ld16(taddr(c, tcol + col), v);
    else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
    if (rows == 128 && col + 16 <= n_valid) {      // whole chunk inside the layer's outputs, every row live: no masks
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], slope * v[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int cc = col + j;
        const float x = fmaxf(v[j], slope * v[j]);
        v[j] = !vr ? 0.f : (cc < n_valid ? x : (cc == n_valid ? 1.f : 0.f));     // bias column: constant 1
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = v[8 * q + j];
      uint4 hh, ll;
      tc::split8(x, hh, ll);
      const long long off = (long long)(2 * ch + q) * 4096;
      *reinterpret_cast<uint4*>(act + off) = hh;
      *reinterpret_cast<uint4*>(act + off + 2048) = ll;
      if (keep) {
        *reinterpret_cast<uint4*>(st + (long long)(2 * ch + q) * st_g) = hh;
        *reinterpret_cast<uint4*>(st + (long long)(2 * ch + q) * st_g + st_lo) = ll;
      }
    }
  }
}

__device__ __forceinline__ void epi_head(EpiCtx& c, const Epi& e) {
  const int h = e.half;
  const bool vr = c.row < c.rows_of(h);
  const int ld = c.pg->lay.ld_mulv;
  float* dst = reinterpret_cast<float*>(c.stash + c.pg->lay.mulv[e.mod]) + (long long)(128 * h + c.row) * ld;
  const int n_cols = e.n_cols, n_valid = e.n_valid, tcol = e.tmem_col;
  for (int ch = c.cpart; ch * 16 < n_cols; ch += c.parts) {
    float v[16];
    __syncwarp();
    tc::tmem_ld16(taddr(c, tcol + ch * 16), v);
    if (vr) {
#pragma unroll
      for (int j = 0; j < 16; ++j) if (ch * 16 + j < n_valid) dst[ch * 16 + j] = v[j];
    }
  }
}

// fusion + reparameterisation + KL for half h; builds decoder inputs [z | c | 1] (cVAE.py:1130-1164, 199)
template <bool RECON>
__device__ void epi_latent(EpiCtx& c, const Epi& e, const float* eps_src) {
  const int rmode = c.template rmode<RECON>();
  const ArchDesc& a = *c.a;
  const Layout& lay = c.pg->lay;
  const int h = e.half, Z = a.Z, M = a.M, rows = c.rows_of(h);
  float* S = c.scratch;
  float* zbuf = reinterpret_cast<float*>(c.stash + lay.zbuf);
  float w[NMB_MAX_MOD];
  if (M > 1 && a.combine == NMB_COMBINE_GPOE) softmax_alpha(c.mb->params + a.alpha_off, M, w);
  const int n = rows * Z;
  const int g_base = (128 * h * Z) / 4;
  // reconstruction: Philox stream 1, counter = tile index (the draws of nmb_ensemble_reconstruct's generic engine)
  float* out_mu = (rmode && c.rt->mu) ? c.rt->mu + (long long)c.row0 * Z : nullptr;
  float* out_lv = (rmode && c.rt->logvar) ? c.rt->logvar + (long long)c.row0 * Z : nullptr;
  for (int g = c.tid; g * 4 < n; g += c.nthr) {
    float nrm[4] = {0.f, 0.f, 0.f, 0.f};
    if (!eps_src && rmode != 1)
      philox_normal4(c.mb->seed, (unsigned long long)c.step, rmode ? 1u : 0u, (uint32_t)(g_base + g), nrm);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int el = g * 4 + j;
      if (el >= n) break;
      const int b = el / Z, z = el - b * Z;
      const int gb = 128 * h + b;
      Fused f;
      if (M == 1) {
        const float* hd = reinterpret_cast<const float*>(c.stash + lay.mulv[0]) + (long long)gb * lay.ld_mulv;
        f.mu = hd[z]; f.lv = hd[Z + z];
      } else {
        float mu[NMB_MAX_MOD], lv[NMB_MAX_MOD];
        for (int m = 0; m < M; ++m) {
          const float* hd = reinterpret_cast<const float*>(c.stash + lay.mulv[m]) + (long long)gb * lay.ld_mulv;
          mu[m] = hd[z]; lv[m] = hd[Z + z];
        }
        f = fuse_forward(mu, lv, M, a.combine, w);
      }
      const float eps = rmode == 1 ? 0.f : (eps_src ? eps_src[gb * Z + z] : nrm[j]);
      S[a.s_mub + gb * Z + z] = f.mu; S[a.s_lvb + gb * Z + z] = f.lv; S[a.s_eps + gb * Z + z] = eps;
      if (out_mu) out_mu[gb * Z + z] = f.mu;
      if (out_lv) out_lv[gb * Z + z] = f.lv;
      zbuf[gb * Z + z] = f.mu + eps * expf(0.5f * f.lv);
      c.kl_acc += -0.5f * (1.f + f.lv - f.mu * f.mu - expf(f.lv));
    }
  }
  bar_n(c.bar_id, c.bar_nthr);
  const int cg = round16(Z + a.C + 1) / 8;
  for (int m = 0; m < M; ++m) {
    const ModDesc& q = a.mod[m];
    unsigned char* st = c.stash + lay.g0[m][h];
    const float* xc = c.template xc_rows<RECON>(m) + (long long)(c.row0 + 128 * h) * q.ldx + q.D;
    const int ldx = q.ldx, C = a.C;
    for (int u = c.tid; u < 128 * cg; u += c.nthr) {
      const int r = u & 127, g = u >> 7;
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cc = 8 * g + j;
        float val = 0.f;
        if (r < rows) {
          if (cc < Z) val = zbuf[(128 * h + r) * Z + cc];
          else if (cc < Z + C) val = xc[(long long)r * ldx + (cc - Z)];
          else if (cc == Z + C) val = 1.f;
        }
        x[j] = val;
      }
      put_planes(st, g, r, x);
      if (m == 0) put_planes(c.smem + h * kActBytes, g, r, x);
    }
  }
}

__device__ __forceinline__ void epi_copy(EpiCtx& c, const Epi& e) {
  const uint4* src = reinterpret_cast<const uint4*>(c.stash + e.src_off);
  uint4* dst = reinterpret_cast<uint4*>(c.smem + e.half * kActBytes);
  const int n = e.src_cg * 256;       // 16-byte units
  for (int u = c.tid; u < n; u += c.nthr) dst[u] = src[u];
}

// Column sums over the 32 rows of a warp: 8 values per lane -> lanes (bit4, bit3, bit2) hold one column each, in
// 4 + 2 + 1 + 1 + 1 shuffles.  Returns the sum of column cj (valid in lanes with (lane & 3) == 0).
__device__ __forceinline__ float warp_colsum8(const float (&qv)[8], int lane, int& cj) {
  float s4[4], s2[2], s1;
  const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float keepv = up16 ? qv[j + 4] : qv[j], send = up16 ? qv[j] : qv[j + 4];
    s4[j] = keepv + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float keepv = up8 ? s4[j + 2] : s4[j], send = up8 ? s4[j] : s4[j + 2];
    s2[j] = keepv + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const float keepv = up4 ? s2[1] : s2[0], send = up4 ? s2[0] : s2[1];
    s1 = keepv + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
  cj = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  return s1;
}

// exp(-logvar_out) per ROI, zero-padded by 8: refreshed whenever logvar_out changes (member start, EK_STEP_END), so
// the reconstruction epilogue needs one coalesced float4 pair per chunk instead of 8 loads + 8 exponentials.
__device__ void build_iv_table(const EpiCtx& c) {
  const ArchDesc& a = *c.a;
  if (a.loss_kind != NMB_LOSS_GAUSS_LL) return;
  for (int m = 0; m < a.M; ++m) {
    const ModDesc& q = a.mod[m];
    float* iv = reinterpret_cast<float*>(c.stash + c.pg->lay.ivtab[m]);
    for (int n = c.tid; n < round4(q.D) + 8; n += c.nthr) iv[n] = n < q.D ? __expf(-__ldcg(c.mb->params + q.lam_off + n)) : 0.f;
  }
}

// x_recon tile: loss terms, d(total)/d(x_recon) planes (ACT[h] or stash), logvar_out gradient partials.
// Gaussian LL per element: -0.5 r^2 e^{-lam} - 0.5 lam - 0.5 log 2pi.  Only the first term is accumulated here; the
// row-independent rest is added once per step in epi_lam.  Rows beyond the minibatch and columns beyond D need no
// masks: their targets, reconstructions (zero operand rows / weight rows) and table entries are exactly zero.
__device__ __forceinline__ void epi_recon(EpiCtx& c, const Epi& e) {
  const ArchDesc& a = *c.a;
  const ModDesc& q = a.mod[e.mod];
  const Layout& lay = c.pg->lay;
  const int h = e.half;
  const bool vr = c.row < c.rows_of(h);
  const int gauss = a.loss_kind == NMB_LOSS_GAUSS_LL;
  const float inv_rows = 1.f / c.rows;
  const float inv_rows_d = inv_rows / q.D;
  const float gscale = gauss ? -inv_rows : -2.f * inv_rows_d;
  unsigned char* st = e.to_act ? c.smem + h * kActBytes : c.stash + e.stash_off;   // d/dx_recon planes
  // ROI targets of this row, lane-major: float4 (col quad) of 32 consecutive rows = 512 contiguous bytes
  const float* __restrict__ xq = c.mt->xlm[e.mod] + ((long long)(c.pos * c.mt->n_half + h) * lay.x_quads[e.mod] * 128 + c.row) * 4;
  const float* ivt = reinterpret_cast<const float*>(c.stash + lay.ivtab[e.mod]);
  float* keep = (c.flags & NMB_TRAIN_KEEP_ACTS) ? c.scratch + q.s_xr + (long long)(128 * h + c.row) * q.ld_xh : nullptr;
  float* lampart = reinterpret_cast<float*>(c.stash + lay.lampart[e.mod]) + (long long)(h * 4 + (c.warp & 3)) * round4(q.D);
  const int n_valid = e.n_valid, col0 = e.col0, tcol = e.tmem_col, n_cols = e.n_cols, D = q.D;
  float ll = 0.f;
  // targets and table entries of chunk k + 1 are in flight while chunk k is processed
  float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xb = xa, ia = xa, ib = xa;
#define NMB_LOAD_X(CH)                                                                                   \
  do {                                                                                                   \
    const int gc_ = col0 + (CH) * 8;                                                                     \
    xa = xb = make_float4(0.f, 0.f, 0.f, 0.f);                                                           \
    if (vr && (CH) * 8 < n_valid) {                                                                      \
      xa = __ldg(reinterpret_cast<const float4*>(xq + (long long)(gc_ >> 2) * 512));                     \
      if (gc_ + 4 < D) xb = __ldg(reinterpret_cast<const float4*>(xq + (long long)((gc_ >> 2) + 1) * 512)); \
    }                                                                                                    \
    if (gauss) {                                                                                         \
      ia = *reinterpret_cast<const float4*>(ivt + gc_);                                                  \
      ib = *reinterpret_cast<const float4*>(ivt + gc_ + 4);                                              \
    }                                                                                                    \
  } while (0)
  if (c.cpart * 8 < n_cols) NMB_LOAD_X(c.cpart);
  float qprev[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) qprev[j] = 0.f;
  int gc_prev = -1, nv_prev = 0;
  for (int ch = c.cpart; ch * 8 < n_cols; ch += c.parts) {
    const int col = ch * 8, gc = col0 + col;
    const int nv = min(max(n_valid - col, 0), 8);
    const float xt[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
    const float iv[8] = {ia.x, ia.y, ia.z, ia.w, ib.x, ib.y, ib.z, ib.w};
    if ((ch + c.parts) * 8 < n_cols) NMB_LOAD_X(ch + c.parts);
    uint32_t raw[8];
    TR3();
    __syncwarp();
    tc::tmem_ld8_issue(taddr(c, tcol + col), raw);
    TR3();
    tc::tmem_ld_wait8(raw);
    TR3();
    // column sums of the previous chunk (zeros on the first pass): unconditional, so that the shuffles sit in the same
    // basic block as this chunk's independent element arithmetic and their latencies overlap it
    int cjp;
    const float s1p = warp_colsum8(qprev, c.lane, cjp);
    float gr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float r = xt[j] - __uint_as_float(raw[j]);
      const float ri = gauss ? r * iv[j] : r;
      const float rri = r * ri;
      ll += rri;
      gr[j] = ri * gscale;
      qprev[j] = rri;
    }
    if (gauss && gc_prev >= 0 && !(c.lane & 3) && cjp < nv_prev) lampart[gc_prev + cjp] = s1p;
    if (!vr) {          // ragged minibatch: keep the invariant "rows beyond the minibatch are zero" explicit
#pragma unroll
      for (int j = 0; j < 8; ++j) gr[j] = 0.f;
    }
    TR3();
    put_planes(st, ch, c.row, gr);
    TR3();
    if (keep && vr && nv > 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j < nv) keep[gc + j] = __uint_as_float(raw[j]);
    }
    gc_prev = gc; nv_prev = nv;
  }
  if (gauss && gc_prev >= 0) {
    int cj;
    const float s1 = warp_colsum8(qprev, c.lane, cj);
    if (!(c.lane & 3) && cj < nv_prev) lampart[gc_prev + cj] = s1;
  }
#undef NMB_LOAD_X
  c.ll_acc += ll * (gauss ? -0.5f * inv_rows : -inv_rows_d);
}

// Forward-only program: x_recon tile (accumulator rows = the 128 rows of this half) -> fp32 rows of the caller's
// [n_rows][D] output (pred_recon, cVAE.py:549-555 / 1198-1208).  Each thread owns one row: 64 contiguous bytes per chunk.
__device__ __forceinline__ void epi_xhat(EpiCtx& c, const Epi& e) {
  const ModDesc& q = c.a->mod[e.mod];
  const int h = e.half, D = q.D;
  const bool vr = c.row < c.rows_of(h);
  float* out = c.rt->xhat[e.mod];
  const int n_valid = e.n_valid, n_cols = e.n_cols, tcol = e.tmem_col;
  float* dst = out ? out + (long long)(c.row0 + 128 * h + c.row) * D + e.col0 : nullptr;
  const bool vec = out && ((reinterpret_cast<unsigned long long>(dst) & 15ull) == 0ull);
  for (int ch = c.cpart; ch * 16 < n_cols; ch += c.parts) {
    const int col = ch * 16;
    if (col >= n_valid) break;
    float v[16];
    __syncwarp();
    tc::tmem_ld16(taddr(c, tcol + col), v);
    if (!vr || !out) continue;
    if (vec && col + 16 <= n_valid) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + col + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) if (col + j < n_valid) dst[col + j] = v[j];
    }
  }
}

// logvar_out: gradient from the per-warp partial sums of r^2 e^{-lam} (both halves complete), Adam, refreshed
// exp(-lam) table; also adds the row-independent part of the Gaussian LL of this step (pre-update values).
__device__ __forceinline__ void epi_lam(EpiCtx& c, const Epi& e) {
  const ModDesc& q = c.a->mod[e.mod];
  const float* part = reinterpret_cast<const float*>(c.stash + c.pg->lay.lampart[e.mod]);
  float* iv = reinterpret_cast<float*>(c.stash + c.pg->lay.ivtab[e.mod]);
  const int ld = round4(q.D);
  const int np = c.rows_h1 > 0 ? 8 : 4;
  const float inv_rows = 1.f / c.rows;
  float llc = 0.f;
  for (int n = c.tid; n < q.D; n += c.nthr) {
    float s = 0.f;
    for (int k = 0; k < np; ++k) s += part[k * ld + n];
    llc += -0.5f * __ldcg(c.mb->params + q.lam_off + n) - 0.5f * kLog2Pi;
    adam_scalar(c, q.lam_off + n, 0.5f - 0.5f * s * inv_rows);       // mean over rows of 0.5 (1 - r^2 e^{-lam})
    iv[n] = __expf(-__ldcg(c.mb->params + q.lam_off + n));
  }
  c.ll_acc += llc;
}

__device__ __forceinline__ void epi_dgrad(EpiCtx& c, const Epi& e) {
  const int h = e.half;
  const int rows = c.rows_of(h);
  const bool vr = c.row < rows;
  unsigned char* act = c.smem + h * kActBytes;
  const int b64 = e.src_cg;        // layout of the stashed activation, see epi_hidden
  const unsigned char* sg = c.stash + e.src_off + (b64 ? (long long)(c.row >> 6) * b64 * 2048 + (c.row & 63) * 16 : (long long)c.row * 16);
  const long long sg_g = b64 ? 2048 : 4096;
  const int nl = c.a->non_linear;
  const int n_cols = e.n_cols, n_mma = e.n_mma, n_valid = e.n_valid, tcol = e.tmem_col;
  // the sign source of chunk k + 1 (hi plane of the stored activation) is in flight while chunk k is processed
  uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
  if (nl && c.cpart * 16 < n_cols) {
    n0 = *reinterpret_cast<const uint4*>(sg + (long long)(2 * c.cpart) * sg_g);
    n1 = *reinterpret_cast<const uint4*>(sg + (long long)(2 * c.cpart + 1) * sg_g);
  }
  for (int ch = c.cpart; ch * 16 < n_cols; ch += c.parts) {
    const int col = ch * 16;
    const uint4 s0 = n0, s1 = n1;
    if (nl && (ch + c.parts) * 16 < n_cols) {
      n0 = *reinterpret_cast<const uint4*>(sg + (long long)(2 * (ch + c.parts)) * sg_g);
      n1 = *reinterpret_cast<const uint4*>(sg + (long long)(2 * (ch + c.parts) + 1) * sg_g);
    }
    float v[16];
    if (col < n_mma) tc::tmem_ld16(taddr(c, tcol + col), v);
    else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
    const uint32_t sw[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    if (nl) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        // stored activation (bf16 hi) > 0  <=>  pre-activation > 0: derivative 1, else the slope
        const float a = __uint_as_float((j & 1) ? (sw[j >> 1] & 0xFFFF0000u) : (sw[j >> 1] << 16));
        v[j] = a > 0.f ? v[j] : kSlope * v[j];
      }
    }
    if (!(rows == 128 && col + 16 <= n_valid)) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (vr && col + j < n_valid) ? v[j] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = v[8 * q + j];
      put_planes(act, 2 * ch + q, c.row, x);
    }
  }
}

__device__ __forceinline__ void epi_dz(EpiCtx& c, const Epi& e) {
  const int h = e.half, Z = c.a->Z;
  const bool vr = c.row < c.rows_of(h);
  float* dz = reinterpret_cast<float*>(c.stash + c.pg->lay.dz) + (long long)(128 * h + c.row) * Z;
  const int n_mma = e.n_mma, tcol = e.tmem_col, acc = e.mod > 0;
  for (int ch = c.cpart; ch * 16 < n_mma; ch += c.parts) {
    float v[16];
    __syncwarp();
    tc::tmem_ld16(taddr(c, tcol + ch * 16), v);
    if (vr) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int cc = ch * 16 + j;
        if (cc < Z) dz[cc] = acc ? dz[cc] + v[j] : v[j];
      }
    }
  }
}

// latent + fusion backward for half h -> d[mu | logvar] planes per modality
__device__ void epi_latent_bwd(EpiCtx& c, const Epi& e) {
  const ArchDesc& a = *c.a;
  const Layout& lay = c.pg->lay;
  const int h = e.half, Z = a.Z, M = a.M, rows = c.rows_of(h);
  const float* S = c.scratch;
  const float* P = c.mb->params;
  const float* dzb = reinterpret_cast<const float*>(c.stash + lay.dz);
  const float inv_rows = 1.f / c.rows;
  float w[NMB_MAX_MOD];
  const bool gpoe = M > 1 && a.combine == NMB_COMBINE_GPOE;
  if (gpoe) softmax_alpha(P + a.alpha_off, M, w);
  for (int el = c.tid; el < rows * Z; el += c.nthr) {
    const int b = el / Z, z = el - b * Z;
    const int gi = (128 * h + b) * Z + z;
    const float mub = S[a.s_mub + gi], lvb = S[a.s_lvb + gi], eps = S[a.s_eps + gi], dz = dzb[gi];
    const float sd = expf(0.5f * lvb);
    const float dmu_bar = dz + M * mub * inv_rows;
    const float dlv_bar = dz * eps * sd * 0.5f + M * (expf(lvb) - 1.f) * 0.5f * inv_rows;
    if (M == 1) {
      float* hd = reinterpret_cast<float*>(c.stash + lay.mulv[0]) + (long long)(128 * h + b) * lay.ld_mulv;
      hd[z] = dmu_bar; hd[Z + z] = dlv_bar;
    } else {
      float mu[NMB_MAX_MOD], lv[NMB_MAX_MOD], dmu[NMB_MAX_MOD], dlv[NMB_MAX_MOD], dw[NMB_MAX_MOD];
      for (int m = 0; m < M; ++m) {
        const float* hd = reinterpret_cast<const float*>(c.stash + lay.mulv[m]) + (long long)(128 * h + b) * lay.ld_mulv;
        mu[m] = hd[z]; lv[m] = hd[Z + z];
      }
      fuse_backward(mu, lv, M, a.combine, w, dmu_bar, dlv_bar, dmu, dlv, gpoe ? dw : nullptr);
      for (int m = 0; m < M; ++m) {
        float* hd = reinterpret_cast<float*>(c.stash + lay.mulv[m]) + (long long)(128 * h + b) * lay.ld_mulv;
        hd[z] = dmu[m]; hd[Z + z] = dlv[m];
        if (gpoe) c.dw_acc[m] += dw[m];
      }
    }
  }
  bar_n(c.bar_id, c.bar_nthr);
  const int cg = round16(2 * Z) / 8;
  for (int m = 0; m < M; ++m) {
    const float* hd0 = reinterpret_cast<const float*>(c.stash + lay.mulv[m]) + (long long)(128 * h) * lay.ld_mulv;
    unsigned char* st = c.stash + lay.dmulv[m][h];
    const int ld = lay.ld_mulv;
    for (int u = c.tid; u < 128 * cg; u += c.nthr) {
      const int r = u & 127, g = u >> 7;
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cc = 8 * g + j;
        x[j] = (r < rows && cc < 2 * Z) ? hd0[(long long)r * ld + cc] : 0.f;
      }
      if (m == 0) put_planes(c.smem + h * kActBytes, g, r, x);
      else put_planes(st, g, r, x);
    }
  }
}

// weight gradient (lane = output row) fused with Adam on the lane-major master state; rewrites the
// BF16 planes of the layer.  Every address is a cursor advanced by a constant stride per chunk.
__device__ __forceinline__ void epi_wgrad(EpiCtx& c, const Epi& e) {
  MemberDev& mb = *c.mb;
  const int o = c.row;
  const bool vo = o < e.p_rows;
  const int p_cols = e.p_cols, n_mma = e.n_mma, col0 = e.col0;
  const bool adam = !(c.flags & NMB_TRAIN_NO_ADAM), wg = c.flags & NMB_TRAIN_WRITE_GRADS;
  const long long R4 = (long long)e.mst_R * 4;                       // floats between consecutive column quads
  const long long m0 = e.mst_off + (long long)o * 4 + (long long)((col0 >> 2) + 2 * c.cpart) * R4;
  float* __restrict__ pp = c.mst_p + m0; float* __restrict__ pm = c.mst_m + m0; float* __restrict__ pv = c.mst_v + m0;
  const long long mstride = 2 * c.parts * R4;
  unsigned char* wp = c.mt->wplanes + e.wp_off + o * 16 + (long long)((col0 >> 3) + c.cpart) * 32 * e.wp_R;
  const long long wstride = (long long)c.parts * 32 * e.wp_R;
  const int lo_off = 16 * e.wp_R;
  uint32_t ta = taddr(c, e.tmem_col + 8 * c.cpart);
  const uint32_t tstride = 8 * c.parts;
  const long long rowbase = e.p_off + (long long)o * e.p_ld;
  const bool live = vo && adam;
  // Software pipeline: the state (p, m, v) of chunk k + 1 is in flight while chunk k is updated and stored; the
  // lines of chunk k + 2 are on their way to L1.
  float4 n0, n1, n2, n3, n4, n5;
  n0 = n1 = n2 = n3 = n4 = n5 = make_float4(0.f, 0.f, 0.f, 0.f);
#define NMB_LOAD_STATE(COL, OFF)                                                                                 \
  do {                                                                                                           \
    n0 = n1 = n2 = n3 = n4 = n5 = make_float4(0.f, 0.f, 0.f, 0.f);                                                \
    if (live && (COL) < p_cols) {                                                                                \
      n0 = ld_stream4(pp + (OFF)); n2 = ld_stream4(pm + (OFF));      \
      n4 = ld_stream4(pv + (OFF));                                                         \
      if ((COL) + 4 < p_cols) {                                                                                  \
        n1 = ld_stream4(pp + (OFF) + R4); n3 = ld_stream4(pm + (OFF) + R4); \
        n5 = ld_stream4(pv + (OFF) + R4);                                                  \
      }                                                                                                          \
    }                                                                                                            \
  } while (0)
  int col = col0 + 8 * c.cpart;
  const int cstep = 8 * c.parts, col_end = col0 + n_mma;
  if (col < col_end) NMB_LOAD_STATE(col, 0);
  for (; col < col_end; col += cstep) {        // 8 columns (one plane group) at a time
    const bool on = vo && col < p_cols;
    const bool full = col + 4 < p_cols;
    float g[8];
    float4 pa = n0, pb = n1, ma = n2, mb4 = n3, va = n4, vb = n5;
    if (col + cstep < col_end) {
      NMB_LOAD_STATE(col + cstep, mstride);
    }
    __syncwarp();
    tc::tmem_ld8(ta, g);
    if (on) {
      if (wg) {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (col + j < p_cols) mb.grads[rowbase + col + j] = g[j];
      }
      if (adam) {
        float x[8];
        x[0] = adam_update(c, ma.x, va.x, pa.x, g[0]); x[1] = adam_update(c, ma.y, va.y, pa.y, g[1]);
        x[2] = adam_update(c, ma.z, va.z, pa.z, g[2]); x[3] = adam_update(c, ma.w, va.w, pa.w, g[3]);
        x[4] = adam_update(c, mb4.x, vb.x, pb.x, g[4]); x[5] = adam_update(c, mb4.y, vb.y, pb.y, g[5]);
        x[6] = adam_update(c, mb4.z, vb.z, pb.z, g[6]); x[7] = adam_update(c, mb4.w, vb.w, pb.w, g[7]);
        st_stream4(pm, ma); st_stream4(pv, va);
        st_stream4(pp, make_float4(x[0], x[1], x[2], x[3]));
        if (full) {
          st_stream4(pm + R4, mb4); st_stream4(pv + R4, vb);
          st_stream4(pp + R4, make_float4(x[4], x[5], x[6], x[7]));
        }
        if (col + 8 > p_cols) {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = (col + j < p_cols) ? x[j] : 0.f;
        }
        uint4 hh, ll;
        tc::split8(x, hh, ll);
        *reinterpret_cast<uint4*>(wp) = hh;
        *reinterpret_cast<uint4*>(wp + lo_off) = ll;
      }
    }
    pp += mstride; pm += mstride; pv += mstride; wp += wstride; ta += tstride;
  }
#undef NMB_LOAD_STATE
}

// transposed weight gradient of decoder_mean_layer: lane = input index i, columns = output rows o.
// Same software pipeline as epi_wgrad: the state of chunk k + 1 is in flight while chunk k is updated.
__device__ __forceinline__ void epi_wgrad_t(EpiCtx& c, const Epi& e) {
  MemberDev& mb = *c.mb;
  const int i = c.row;
  const bool vi = i < e.p_cols;
  const int p_ld = e.p_ld, p_rows = e.p_rows, n_mma = e.n_mma, col0 = e.col0;
  const long long R4 = (long long)e.mst_R * 4;
  // the layer's planes are stored transposed (rows = input index): this lane's 8 output rows are one 16-byte element
  unsigned char* wp = c.mt->wplanes + e.wp_off + i * 16;
  const long long grp_bytes = 32LL * e.wp_R;
  const int lo_off = 16 * e.wp_R;
  const bool adam = !(c.flags & NMB_TRAIN_NO_ADAM), wg = c.flags & NMB_TRAIN_WRITE_GRADS;
  const long long colbase = e.p_off + i;
  const long long m0 = e.mst_off + (long long)i * 4 + (long long)((col0 >> 2) + 2 * c.cpart) * R4;
  float* __restrict__ pp = c.mst_p + m0; float* __restrict__ pm = c.mst_m + m0; float* __restrict__ pv = c.mst_v + m0;
  const long long mstride = 2 * c.parts * R4;
  uint32_t ta = taddr(c, e.tmem_col + 8 * c.cpart);
  const uint32_t tstride = 8 * c.parts;
  const bool live = vi && adam;
  float4 n0, n1, n2, n3, n4, n5;
  n0 = n1 = n2 = n3 = n4 = n5 = make_float4(0.f, 0.f, 0.f, 0.f);
#define NMB_LOAD_STATE_T(O0, OFF)                                                                                \
  do {                                                                                                           \
    n0 = n1 = n2 = n3 = n4 = n5 = make_float4(0.f, 0.f, 0.f, 0.f);                                                \
    if (live && (O0) < p_rows) {                                                                                 \
      n0 = ld_stream4(pp + (OFF)); n2 = ld_stream4(pm + (OFF)); n4 = ld_stream4(pv + (OFF));                     \
      if ((O0) + 4 < p_rows) {                                                                                   \
        n1 = ld_stream4(pp + (OFF) + R4); n3 = ld_stream4(pm + (OFF) + R4); n5 = ld_stream4(pv + (OFF) + R4);    \
      }                                                                                                          \
    }                                                                                                            \
  } while (0)
  int o0 = col0 + 8 * c.cpart;
  const int ostep = 8 * c.parts, o_end = col0 + n_mma;
  if (o0 < o_end) NMB_LOAD_STATE_T(o0, 0);
  for (; o0 < o_end; o0 += ostep) {
    const bool on = vi && o0 < p_rows;
    const bool full = o0 + 4 < p_rows;
    float g[8];
    float4 pa = n0, pb = n1, ma = n2, mb4 = n3, va = n4, vb = n5;
    if (o0 + ostep < o_end) NMB_LOAD_STATE_T(o0 + ostep, mstride);
    __syncwarp();
    tc::tmem_ld8(ta, g);
    if (on) {
      if (wg) {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (o0 + j < p_rows) mb.grads[colbase + (long long)(o0 + j) * p_ld] = g[j];
      }
      if (adam) {
        float x[8];
        x[0] = adam_update(c, ma.x, va.x, pa.x, g[0]); x[1] = adam_update(c, ma.y, va.y, pa.y, g[1]);
        x[2] = adam_update(c, ma.z, va.z, pa.z, g[2]); x[3] = adam_update(c, ma.w, va.w, pa.w, g[3]);
        x[4] = adam_update(c, mb4.x, vb.x, pb.x, g[4]); x[5] = adam_update(c, mb4.y, vb.y, pb.y, g[5]);
        x[6] = adam_update(c, mb4.z, vb.z, pb.z, g[6]); x[7] = adam_update(c, mb4.w, vb.w, pb.w, g[7]);
        st_stream4(pm, ma); st_stream4(pv, va);
        st_stream4(pp, make_float4(x[0], x[1], x[2], x[3]));
        if (full) {
          st_stream4(pm + R4, mb4); st_stream4(pv + R4, vb);
          st_stream4(pp + R4, make_float4(x[4], x[5], x[6], x[7]));
        }
        if (o0 + 8 > p_rows) {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = (o0 + j < p_rows) ? x[j] : 0.f;
        }
        uint4 hh, ll;
        tc::split8(x, hh, ll);
        unsigned char* pb0 = wp + (long long)(o0 >> 3) * grp_bytes;
        *reinterpret_cast<uint4*>(pb0) = hh;
        *reinterpret_cast<uint4*>(pb0 + lo_off) = ll;
      }
    }
    pp += mstride; pm += mstride; pv += mstride; ta += tstride;
  }
#undef NMB_LOAD_STATE_T
}

// ---- one modality, Z <= 16: head accumulator -> reparameterisation -> decoder input, all in registers --------
// One thread per minibatch row (the 4 warps of the group that own the TMEM lanes); cVAE.py:1130-1139, 199.
__device__ __forceinline__ float bf16pair_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16pair_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// Pre-phase (before the accumulator barrier is waited on, so it overlaps the head GEMM): the Philox draws of this
// half, spread over the whole group, written to the eps array the backward pass reads anyway.
template <bool RECON>
__device__ void epi_head_latent_pre(EpiCtx& c, const Epi& e, const float* eps_src) {
  const int rmode = c.template rmode<RECON>();
  {   // this row's decoder-input template planes start their way to L1 while the head GEMM finishes (no registers held)
    const Layout& lay = c.pg->lay;
    const unsigned char* tp = c.template cplanes0<RECON>() + (long long)(c.pos * c.template n_half<RECON>() + e.half) * lay.c_cg * 4096 + c.row * 16;
    for (int g = 0; g < lay.c_cg; ++g) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(tp + (long long)g * 4096));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(tp + (long long)g * 4096 + 2048));
    }
  }
  if (eps_src || rmode == 1) return;
  const ArchDesc& a = *c.a;
  const int h = e.half, Z = a.Z, n = c.rows_of(h) * Z;
  float* eps = c.scratch + a.s_eps + 128 * h * Z;
  const int g_base = (128 * h * Z) / 4;
  for (int g = c.tid; g * 4 < n; g += c.nthr) {
    float nrm[4];
    philox_normal4(c.mb->seed, (unsigned long long)c.step, rmode ? 1u : 0u, (uint32_t)(g_base + g), nrm);
#pragma unroll
    for (int j = 0; j < 4; ++j) if (g * 4 + j < n) eps[g * 4 + j] = nrm[j];
  }
  bar_n(c.bar_id, c.bar_nthr);
}

template <bool RECON>
__device__ void epi_head_latent(EpiCtx& c, const Epi& e, const float* eps_src) {
  const int rmode = c.template rmode<RECON>();
  
  const ArchDesc& a = *c.a;
  const Layout& lay = c.pg->lay;
  const int h = e.half, Z = a.Z, C = a.C, rows = c.rows_of(h);
  const bool vr = c.row < rows;
  const int gb = 128 * h + c.row;
  float mu[16], lv[16], zz[16];
  float* S = c.scratch;
  const uint32_t e0 = (uint32_t)gb * (uint32_t)Z;     // first element of this row in the [rows x Z] eps stream
  const float* __restrict__ epsrow = (eps_src ? eps_src : S + a.s_eps) + e0;
  // all draws of the row first (a store to the scratch array could alias a later load: the compiler keeps program
  // order, which exposed one L2 round trip per latent element); zz[] holds eps until the arithmetic
#pragma unroll
  for (int z = 0; z < 16; ++z) zz[z] = (z < Z && vr && rmode != 1) ? epsrow[z] : 0.f;    // mean decode: eps = 0
  TR3();
  tc::tmem_ld16(taddr(c, e.tmem_col), mu);            // columns 0 .. 15: mu[0 .. Z)
  tc::tmem_ld16(taddr(c, e.tmem_col + Z), lv);        // columns Z .. Z + 15: logvar[0 .. Z)
  TR3();
  float kl = 0.f;
  const long long s_mub = a.s_mub, s_lvb = a.s_lvb, s_eps = a.s_eps;
  if (rmode) {            // pred_latent outputs of a reconstruction launch (rows of this tile)
    float* om = c.rt->mu ? c.rt->mu + ((long long)c.row0 * Z + e0) : nullptr;
    float* ol = c.rt->logvar ? c.rt->logvar + ((long long)c.row0 * Z + e0) : nullptr;
#pragma unroll
    for (int z = 0; z < 16; ++z) {
      if (z < Z && vr) {
        if (om) om[z] = mu[z];
        if (ol) ol[z] = lv[z];
      }
    }
  }
#pragma unroll
  for (int z = 0; z < 16; ++z) {
    if (z < Z && vr) {
      const float m_ = mu[z], l_ = lv[z], eps = zz[z];
      S[s_mub + e0 + z] = m_; S[s_lvb + e0 + z] = l_;
      if (eps_src) S[s_eps + e0 + z] = eps;
      zz[z] = m_ + eps * __expf(0.5f * l_);
      kl += -0.5f * (1.f + l_ - m_ * m_ - __expf(l_));
    }
  }
  c.kl_acc += kl;
  TR3();
  // [z | c | 1]: the covariate part comes from the dataset's template block (coalesced 16-byte reads)
  const unsigned char* tp = c.template cplanes0<RECON>() + (long long)(c.pos * c.template n_half<RECON>() + h) * lay.c_cg * 4096 + c.row * 16;
  unsigned char* act = c.smem + h * kActBytes + c.row * 16;
  unsigned char* st = c.stash + lay.g0[0][h] + c.row * 16;
  const int zg = (Z + 7) >> 3;                          // groups that contain z columns (<= 2)
  for (int g = 0; g < lay.c_cg; ++g) {
    uint4 hi = __ldg(reinterpret_cast<const uint4*>(tp + (long long)g * 4096));          // read-only dataset planes
    uint4 lo = __ldg(reinterpret_cast<const uint4*>(tp + (long long)g * 4096 + 2048));
    if (g < zg) {
      const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = (j & 1) ? bf16pair_hi(hw[j >> 1]) + bf16pair_hi(lw[j >> 1]) : bf16pair_lo(hw[j >> 1]) + bf16pair_lo(lw[j >> 1]);
        const int cc = 8 * g + j;
        x[j] = cc < Z ? (g == 0 ? zz[j] : zz[8 + j]) : t;
      }
      tc::split8(x, hi, lo);
    }
    *reinterpret_cast<uint4*>(act + (long long)g * 4096) = hi;
    *reinterpret_cast<uint4*>(act + (long long)g * 4096 + 2048) = lo;
    *reinterpret_cast<uint4*>(st + (long long)g * 4096) = hi;
    *reinterpret_cast<uint4*>(st + (long long)g * 4096 + 2048) = lo;
    TR3();
  }
  (void)C;
}

// ---- one modality, Z <= 16: d/dz accumulator -> latent backward -> d[mu | logvar] planes in ACT[h] ------------
__device__ void epi_dz_latent_bwd(EpiCtx& c, const Epi& e) {
  const ArchDesc& a = *c.a;
  const int h = e.half, Z = a.Z, rows = c.rows_of(h);
  const bool vr = c.row < rows;
  const int gb = 128 * h + c.row;
  const float* S = c.scratch;
  const float inv_rows = 1.f / c.rows;
  const long long s_mub = a.s_mub, s_lvb = a.s_lvb, s_eps = a.s_eps;
  float dz[16], dmu[16], dlv[16];
  tc::tmem_ld16(taddr(c, e.tmem_col), dz);
  // phase 1: every global load and all arithmetic, statically indexed registers only
#pragma unroll
  for (int z = 0; z < 16; ++z) {
    dmu[z] = 0.f; dlv[z] = 0.f;
    if (z < Z && vr) {
      const int gi = gb * Z + z;
      const float mub = S[s_mub + gi], lvb = S[s_lvb + gi], eps = S[s_eps + gi];
      const float sd = __expf(0.5f * lvb);
      dmu[z] = dz[z] + mub * inv_rows;                                                   // d/dmu  (M = 1)
      dlv[z] = dz[z] * eps * sd * 0.5f + (__expf(lvb) - 1.f) * 0.5f * inv_rows;           // d/dlogvar
    }
  }
  // phase 2: planes of [d/dmu (Z) | d/dlogvar (Z) | 0] in ACT[h].  d/dmu sits at static columns; d/dlogvar starts at the
  // run-time column Z, so its BF16 hi / lo elements are placed with 2-byte shared-memory stores over a zero fill
  // (same thread, same row: program order).
  unsigned char* act = c.smem + h * kActBytes + c.row * 16;
  const int cg = round16(2 * Z) / 8;
  {
    float x[8];
    uint4 hh, ll;
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = dmu[j];
    tc::split8(x, hh, ll);
    *reinterpret_cast<uint4*>(act) = hh; *reinterpret_cast<uint4*>(act + 2048) = ll;
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = dmu[8 + j];
    tc::split8(x, hh, ll);
    if (cg > 1) { *reinterpret_cast<uint4*>(act + 4096) = hh; *reinterpret_cast<uint4*>(act + 4096 + 2048) = ll; }
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int g = 2; g < cg; ++g) { *reinterpret_cast<uint4*>(act + g * 4096) = zero; *reinterpret_cast<uint4*>(act + g * 4096 + 2048) = zero; }
  }
#pragma unroll
  for (int z = 0; z < 16; z += 2) {
    if (z < Z) {
      uint32_t hi, lo;
      tc::split2(dlv[z], dlv[z + 1], hi, lo);
      const int c0 = Z + z, c1 = Z + z + 1;
      unsigned char* p0 = act + (c0 >> 3) * 4096 + (c0 & 7) * 2;
      *reinterpret_cast<unsigned short*>(p0) = (unsigned short)(hi & 0xFFFFu);
      *reinterpret_cast<unsigned short*>(p0 + 2048) = (unsigned short)(lo & 0xFFFFu);
      if (z + 1 < Z) {
        unsigned char* p1 = act + (c1 >> 3) * 4096 + (c1 & 7) * 2;
        *reinterpret_cast<unsigned short*>(p1) = (unsigned short)(hi >> 16);
        *reinterpret_cast<unsigned short*>(p1 + 2048) = (unsigned short)(lo >> 16);
      }
    }
  }
}

// While the optimiser group waits for a weight-gradient accumulator it pulls that item's Adam state towards L2
// (the per-slot master buffers of 148 resident members exceed the L2, so the first touch is an HBM miss).
__device__ __forceinline__ void prefetch_adam_state(const EpiCtx& c, const Epi& e) {
  if (c.flags & NMB_TRAIN_NO_ADAM) return;
  const bool t_ = e.kind == EK_WGRAD_T;
  const int lanes = t_ ? e.p_cols : e.p_rows, other = t_ ? e.p_rows : e.p_cols;
  if (c.row >= lanes) return;
  const int q0 = e.col0 >> 2, q1 = min((e.col0 + e.n_mma + 3) >> 2, (other + 3) >> 2);
  const long long base = e.mst_off + (long long)c.row * 4;
  for (int qd = q0; qd < q1; ++qd) {
    const long long mi = base + (long long)qd * e.mst_R * 4;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c.mst_p + mi));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c.mst_m + mi));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(c.mst_v + mi));
  }
}

template <bool RECON>
__device__ void epi_step_end(EpiCtx& c, float* loss_out) {
  const int rmode = c.template rmode<RECON>();
  const ArchDesc& a = *c.a;
  const int M = a.M;
  if (rmode) { c.kl_acc = 0.f; c.ll_acc = 0.f; return; }
  // both loss sums in one pass over the barrier pair (red[] holds 12 + 12 partials)
  float kl, ll;
  {
    const float a0 = warp_sum(c.kl_acc), a1 = warp_sum(c.ll_acc);      // (the item started with a rendezvous: red[] is free)
    if (c.lane == 0) { c.ctl->red[c.warp] = a0; c.ctl->red[kEpiWarps + c.warp] = a1; }
    bar_n(4, kEpiWarps * 32);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < kEpiWarps; ++i) { s0 += c.ctl->red[i]; s1 += c.ctl->red[kEpiWarps + i]; }
    kl = s0 / c.rows; ll = s1;
  }
  if (loss_out && c.tid == 0) { loss_out[0] = M * kl - ll; loss_out[1] = M * kl; loss_out[2] = ll; }
  if (M > 1 && a.combine == NMB_COMBINE_GPOE) {
    float w[NMB_MAX_MOD], dw_tot[NMB_MAX_MOD];
    softmax_alpha(c.mb->params + a.alpha_off, M, w);
    for (int m = 0; m < M; ++m) dw_tot[m] = block_sum_epi(c, c.dw_acc[m]);
    bar_n(c.bar_id, c.bar_nthr);
    if (c.tid == 0) {
      float dot = 0.f;
      for (int m = 0; m < M; ++m) dot += w[m] * dw_tot[m];
      for (int m = 0; m < M; ++m) adam_scalar(c, a.alpha_off + m, w[m] * (dw_tot[m] - dot));
    }
  }
  c.kl_acc = 0.f; c.ll_acc = 0.f;
  for (int m = 0; m < M; ++m) c.dw_acc[m] = 0.f;
}

// Worker set of an item: the item's group (128 threads, one warp per TMEM lane quadrant) or all epilogue warps.
__device__ __forceinline__ void set_workers(EpiCtx& c, bool all) {
  const int lw = c.warp % kGroupWarps;
  c.cpart = 0; c.parts = 1;
  if (all) {
    c.tid = c.warp * 32 + c.lane; c.nthr = kEpiWarps * 32; c.bar_id = 4; c.bar_nthr = kEpiWarps * 32;
  } else {
    c.tid = lw * 32 + c.lane; c.nthr = kGroupThreads; c.bar_id = 1 + c.grp; c.bar_nthr = kGroupThreads;
  }
}

// The epilogue runs as THREE groups of 4 warps: one per 128-row half of the minibatch (activations, losses,
// data gradients: two independent dependency chains whose latencies overlap) and the optimiser group, which
// consumes the weight-gradient accumulators (Adam + new weight planes) off both chains.  Every group walks the
// item list and executes its own items; EK_STEP_END is the only rendezvous.
template <bool RECON>
__device__ void epilogue_role(const LaunchP& L, int ai, int mi, EpiCtx& c, const RowSet& rs, uint32_t& acc_par) {
  const int rmode = c.template rmode<RECON>();
  const TrainLaunch& t = L.t;
  const ProgramDev& pg = *c.pg;
  MemberDev& mb = *c.mb;
  c.kl_acc = 0.f; c.ll_acc = 0.f;
  float dw_acc[NMB_MAX_MOD];
  for (int m = 0; m < NMB_MAX_MOD; ++m) dw_acc[m] = 0.f;
  c.dw_acc = dw_acc;
  set_workers(c, true);
  if (!rmode) build_iv_table(c);   // weight planes and the lane-major master state were prepared by tcp_prepare_kernel
  __threadfence();
  fence_async_all();
  bar_n(4, kEpiWarps * 32);
  if (c.tid == 0) {
#pragma unroll
    for (int g = 0; g < kGroups; ++g) st_release(&c.ctl->epi_done[g], 1u);
  }
  const long long s0 = c.ctl->chunk_s0;
  const int i0 = c.ctl->chunk_i0, n_chunk = c.ctl->chunk_n;
  const int n_epis = pg.n_epis;
  const Epi* __restrict__ epis = pg.epis;
  const bool in_params = L.ep_cnt[ai] > 0;      // item table in the kernel parameters (constant bank)
  const int ep0 = L.ep_off[ai];
  const bool pub = (c.warp % kGroupWarps) == 0 && c.lane == 0;     // the thread that publishes its group's counter
  for (long long i = 0; i < n_chunk; ++i) {
    const long long s = s0 + i;
    const StepVars sv = step_vars(rs, s, i, n_epis);
    c.rows = sv.rows; c.rows_h0 = sv.rows_h[0]; c.rows_h1 = sv.rows_h[1]; c.row0 = sv.row0; c.pos = sv.pos;
    c.step = s;
    const float* eps = nullptr;
    float* lo = nullptr;
    if (rmode) {           // s = tile index; injected draws [n_rows][Z] of the caller, rows of this tile
      if (rmode == 2 && c.rt->eps) eps = c.rt->eps + (long long)sv.row0 * c.a->Z;
      c.step_size = 0.f; c.inv_bc2 = 1.f;
    } else {
      const double tt = (double)(s + 1);
      const float lr = mb.lr_steps ? mb.lr_steps[s] : mb.lr;
      c.step_size = (float)((double)lr / (1.0 - pow((double)mb.beta1, tt)));
      c.inv_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)mb.beta2, tt)));
      eps = t.eps_override ? t.eps_override + ((long long)mi * t.stride_steps + i0 + i) * mb.batch * c.a->Z : nullptr;
      lo = t.loss_out ? t.loss_out + ((long long)mi * t.stride_steps + i0 + i) * 3 : nullptr;
    }
    // Full minibatch + compact table: every group jumps from own item to own item (host-linked list); otherwise it
    // walks the whole list and skips what it does not own.
    const bool jump = in_params && c.rows_h1 > 0 && L.jump;
    const bool tr_step = g_trace && blockIdx.x == 0 && i0 + i == g_trace_step;     // one global read per step, not per item
    for (int k = jump ? (int)L.ep_first[ai][c.grp] : 0; k < n_epis; ) {
      Epi e;
      int k_next = k + 1;
      if (in_params) {
        const EpiP& q = L.epis_p[ep0 + k];
        if (jump) {
          acc_par ^= (uint32_t)q.flip_before[c.grp] << 2;
          k_next = q.next_own[c.grp] == 255 ? n_epis : (int)q.next_own[c.grp];
        }
        e = from_epip(q);
      } else {
        e = epis[k];
        if (c.lane == 0 && k + 2 < n_epis) asm volatile("prefetch.global.L1 [%0];" ::"l"(epis + k + 2));
      }
      const int k_cur = k;
      k = k_next;
      const bool all = e.kind == EK_STEP_END;
      const bool optim = e.kind == EK_WGRAD || e.kind == EK_WGRAD_T;
      if (!jump) {
        const int owner = (all || e.split_all) ? c.grp : (optim ? 2 : e.half);
        if (owner != c.grp || (e.half == 1 && c.rows_h1 == 0)) {
          if (optim) acc_par ^= 1u << e.buf;       // keep this group's view of the accumulator barrier phases in step
          continue;
        }
      }
      set_workers(c, all);
      if (e.split_all) {
        c.cpart = c.grp; c.parts = kGroups;
        // an idle activation group may be far ahead of the optimiser: mbarrier parity only disambiguates one phase
        if (c.grp != 2 && e.wait_optim) wait_epi(&c.ctl->epi_done[2], sv.base + (uint32_t)e.wait_optim);
      }
      const bool tr = tr_step && pub && (!all || c.grp == 0);
#ifdef NMB_TCP_FINE_TRACE
      // loop-level cycle stamps of group 0's publishing thread: [decoded | before accumulator wait | after it | item done |
      // fenced | group barrier | published]
      unsigned long long* lp = (tr && c.grp == 0) ? g_trace + 2048 + 512 + 8 * k_cur : nullptr;
#define LPS(j) do { if (lp) lp[j] = (unsigned long long)clock64(); } while (0)
#else
#define LPS(j) do { } while (0)
#endif
      LPS(0);
      const int tbase = c.grp == 0 ? 0 : 3 * n_epis * c.grp + 5 * pg.n_steps;   // groups 1, 2 stamp after the MMA / producer records
      if (tr) g_trace[tbase + 3 * k_cur] = gtime();
      if (e.kind == EK_LAM) {        // both halves' reconstruction items (and their stash fences) are behind these counters
        wait_epi(&c.ctl->epi_done[0], sv.base + (uint32_t)e.n_valid);
        if (c.rows_h1 > 0) wait_epi(&c.ctl->epi_done[1], sv.base + (uint32_t)e.n_cols);
      }
      if (all) bar_n(4, kEpiWarps * 32);           // every group has finished everything before this item
      if (e.kind == EK_HEAD_LATENT) epi_head_latent_pre<RECON>(c, e, eps);
      // every stash block of this half's forward pass has been written by now and this item writes none: the
      // generic -> async proxy publication (MEMBAR.GPU + proxy fence) runs while the group waits for its accumulator
      if (e.kind == EK_RECON && e.src_cg) { __threadfence(); fence_async_all(); }
      if (e.kind == EK_RECON) {      // first chunk's targets and exp(-logvar_out) entries towards L1 during the wait
        const Layout& lay = pg.lay;
        const float* xq = c.mt->xlm[e.mod] + ((long long)(c.pos * c.mt->n_half + e.half) * lay.x_quads[e.mod] * 128 + c.row) * 4
                          + (long long)(e.col0 >> 2) * 512;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(xq));
        if (e.col0 + 4 < c.a->mod[e.mod].D) asm volatile("prefetch.global.L1 [%0];" ::"l"(xq + 512));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(c.stash + lay.ivtab[e.mod] + (long long)e.col0 * 4));
      }
      if (optim) prefetch_adam_state(c, e);
      LPS(1);
      if (e.buf >= 0) {
        tc::mbar_wait(&c.ctl->accbar[e.buf], (acc_par >> e.buf) & 1u);
        acc_par ^= 1u << e.buf;
        tc::fence_after();
      }
      LPS(2);
      if (tr) g_trace[tbase + 3 * k_cur + 1] = gtime();
#ifdef NMB_TCP_FINE_TRACE
      c.tr_ptr = nullptr;
      if (tr && c.grp == 0 && e.half == 0) {
        if (e.kind == EK_HEAD_LATENT) c.tr_ptr = g_trace + 2048;
        if (e.kind == EK_RECON && e.col0 == 0) c.tr_ptr = g_trace + 2048 + 64;
      }
#endif
      // proxy fences: ACT[h] (shared memory, read by the next MMAs) per item; global data read by the TMA
      // (stash blocks, weight planes) only at EK_FENCE / EK_STEP_END, so the stores drain in the background
      int fence = 0;              // 1 = shared memory, 2 = everything
      switch (e.kind) {
        case EK_HIDDEN: epi_hidden<RECON>(c, e); fence = 1; break;
        case EK_HEAD: epi_head(c, e); break;
        case EK_LATENT: epi_latent<RECON>(c, e, eps); fence = 1; break;
        case EK_COPY: epi_copy(c, e); fence = 1; break;
        case EK_RECON: epi_recon(c, e); fence = e.to_act ? 1 : 0; break;
        case EK_DGRAD: epi_dgrad(c, e); fence = 1; break;
        case EK_DZ: epi_dz(c, e); break;
        case EK_LATENT_BWD: epi_latent_bwd(c, e); fence = 1; break;
        case EK_WGRAD: epi_wgrad(c, e); break;
        case EK_WGRAD_T: epi_wgrad_t(c, e); break;
        case EK_FENCE: fence = e.src_cg ? 0 : 2; break;      // src_cg: already published before the last forward item
        case EK_LAM: epi_lam(c, e); break;
        case EK_HEAD_LATENT: epi_head_latent<RECON>(c, e, eps); fence = 1; break;
        case EK_DZ_LATENT_BWD: epi_dz_latent_bwd(c, e); fence = 1; break;
        case EK_XHAT: if (RECON) epi_xhat(c, e); break;
        default: epi_step_end<RECON>(c, lo); fence = rmode ? 0 : 2; break;
      }
      LPS(3);
      if (e.buf >= 0) tc::fence_before();
      if (fence == 1) fence_async_smem();
      else if (fence == 2) { __threadfence(); fence_async_all(); }
      LPS(4);
      bar_n(1 + c.grp, kGroupThreads);
      LPS(5);
      if (pub) st_release(&c.ctl->epi_done[c.grp], sv.base + (uint32_t)k_cur + 1u);
      LPS(6);
      if (tr) g_trace[tbase + 3 * k_cur + 2] = gtime();
    }
  }
}

// RECON = false: the training kernel (recon_mode is the compile-time constant 0 in every role);
// RECON = true: the forward-only instantiation used by nmb_ensemble_reconstruct.
template <bool RECON>
__global__ void __launch_bounds__(kThreadsP, 1) train_tcp_kernel(const __grid_constant__ LaunchP L) {
  extern __shared__ __align__(1024) unsigned char smem[];
  Ctrl* ctl = reinterpret_cast<Ctrl*>(smem + kSmemCtrl);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TrainLaunch& t = L.t;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) { tc::mbar_init(&ctl->full[i], 1); tc::mbar_init(&ctl->empty[i], 1); }
    for (int i = 0; i < 6; ++i) tc::mbar_init(&ctl->accbar[i], 1);
  }
  if (warp == kEpiWarps) tc::tmem_alloc(&ctl->tmem, 512);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = ctl->tmem;
  uint32_t seq = 0, acc_par = 0;
  unsigned char* stash = L.stash + (long long)blockIdx.x * L.stash_bytes;
  // Work items = (member, chunk of consecutive minibatch steps), dealt longest-member-first, all chunks 0 before
  // all chunks 1 ...  A chunk waits (thread 0, acquire) until the member's previous chunk -- possibly on another
  // SM -- has published its last step: finer items than whole members keep the tail of the launch short.
  constexpr bool recon = RECON;
  const int n_items = recon ? L.n_rwork : t.n_members * L.n_chunks;
  const bool dynamic = (int)gridDim.x < n_items;
  bool first = true;
  for (;;) {
    if (threadIdx.x == 0) {
      int it = n_items;
      if (dynamic) it = atomicAdd(t.work_counter, 1);
      else if (first) it = blockIdx.x;
      int mi = t.n_members;
      if (it < n_items && recon) {          // forward-only launch: (member, range of 256-row tiles), no cross-item order
        const ReconWork rw = L.rwork[it];
        mi = rw.member;
        ctl->chunk_i0 = rw.tile0; ctl->chunk_n = rw.n_tiles; ctl->chunk_s0 = rw.tile0; ctl->rt_idx = rw.rt;
      } else if (it < n_items) {
        const int chunk = it / t.n_members;
        mi = dynamic ? t.order[it - chunk * t.n_members] : it;
        MemberDev& m = t.members[mi];
        const long long ns = member_steps(t, m);
        const long long i0 = ns * chunk / L.n_chunks, i1 = ns * (chunk + 1) / L.n_chunks;
        const long long need = m.launch_base + i0;
        if (chunk > 0) {
          const long long t0 = clock64();
          long long have;
          for (;;) {
            asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(have) : "l"(&m.steps_done) : "memory");
            if (have >= need) break;
            __nanosleep(200);
            if (g_spin_budget > 0 && clock64() - t0 > g_spin_budget) __trap();
          }
        }
        ctl->chunk_i0 = (int)i0; ctl->chunk_n = (int)(i1 - i0); ctl->chunk_s0 = need;
      }
      ctl->member = mi;
      if (mi < t.n_members) {
        if (recon) { ctl->rs.n_rows = L.rtc[ctl->rt_idx].n_rows; ctl->rs.batch = 256; ctl->rs.n_half = 2; }
        else { ctl->rs.n_rows = t.members[mi].n_rows; ctl->rs.batch = t.members[mi].batch; ctl->rs.n_half = L.mtc[mi].n_half; }
      }
      for (int g = 0; g < kGroups; ++g) ctl->epi_done[g] = 0;
    }
    first = false;
    __syncthreads();
    const int mi = ctl->member;
    if (mi >= t.n_members) break;
    MemberDev& mb = t.members[mi];
    const ProgramDev& pg = L.progs[mb.arch_idx];
    const MemberTc& mt = L.mtc[mi];
    const ReconTc* rt = recon ? L.rtc + ctl->rt_idx : nullptr;
    const RowSet& rs = ctl->rs;
    if (ctl->chunk_n > 0) {
      if (warp < kEpiWarps) {
        EpiCtx c;
        c.a = &t.archs[mb.arch_idx]; c.pg = &pg; c.mb = &mb; c.mt = &mt;
        c.recon = recon ? L.recon_mode : 0; c.rt = rt;
        c.smem = smem; c.stash = stash; c.scratch = t.scratch + (long long)blockIdx.x * t.slot_floats; c.ctl = ctl;
        c.mst_p = L.master + (long long)mi * 3 * L.master_floats;        // per member, persistent across work items
        c.mst_m = c.mst_p + L.master_floats; c.mst_v = c.mst_m + L.master_floats;
        c.tmem = tmem; c.warp = warp; c.lane = lane; c.row = ((warp & 3) << 5) + lane;
        c.grp = warp / kGroupWarps; c.flags = t.flags;
        c.b1 = mb.beta1; c.b2 = mb.beta2; c.aeps = mb.adam_eps;
        epilogue_role<RECON>(L, mb.arch_idx, mi, c, rs, acc_par);
      } else if (warp == kEpiWarps) {
        const int ai = __shfl_sync(0xffffffffu, mb.arch_idx, 0);
        mma_role(L, ai, rs, smem, ctl, tmem, seq, pg.n_epis);
      } else {
        if (lane == 0) producer_role(t, pg, rs, mt, recon ? rt->xplanes : mt.xplanes, recon, stash, smem, ctl, seq);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && !recon) {
      const int spe = (mb.n_rows + mb.batch - 1) / mb.batch;
      const long long done = ctl->chunk_s0 + ctl->chunk_n;
      if (ctl->chunk_n > 0) {
        mb.last_rows = min(mb.batch, mb.n_rows - (int)((done - 1) % spe) * mb.batch);
        mb.last_slot = blockIdx.x;
      }
      __threadfence();       // everything this CTA wrote for the member is visible before the step count is
      asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(&mb.steps_done), "l"(done) : "memory");
    }
    __syncthreads();
  }
  tc::fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tc::tmem_free(tmem, 512);
}

// ---- dataset planes: packed fp32 rows [x | c | 1] -> 128-row canonical blocks per (minibatch, half) ----
__global__ void __launch_bounds__(256) xprep_kernel(const XPrepItem* items, int n_items) {
  for (int it = blockIdx.y; it < n_items; it += gridDim.y) {
    const XPrepItem x = items[it];
    const int spe = (x.n_rows + x.batch - 1) / x.batch;
    const int n_blocks = spe * x.n_half;
    for (int b = blockIdx.x; b < n_blocks; b += gridDim.x) {
      const int pos = b / x.n_half, h = b - pos * x.n_half;
      const int r0 = pos * x.batch + 128 * h;
      int rows = min(x.batch - 128 * h, x.n_rows - r0);
      rows = rows > 128 ? 128 : rows;
      unsigned char* blk = x.out + (long long)b * x.cg * 4096;
      for (int u = threadIdx.x; u < 128 * x.cg; u += 256) {
        const int r = u & 127, g = u >> 7;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
        if (r < rows) {
          const float* src = x.xc + (long long)(r0 + r) * x.ldx + 8 * g;
          const int nv = x.k_valid - 8 * g;
#pragma unroll
          for (int j = 0; j < 8; ++j) if (j < nv) v[j] = src[j];
        }
        put_planes(blk, g, r, v);
      }
      // decoder-input template [0 (Z) | c | 1 | 0]
      unsigned char* cb = x.cplanes + (long long)b * x.c_cg * 4096;
      for (int u = threadIdx.x; u < 128 * x.c_cg; u += 256) {
        const int r = u & 127, g = u >> 7;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int cc = 8 * g + j;
          float val = 0.f;
          if (r < rows) {
            if (cc >= x.z && cc < x.z + x.c_dim) val = x.xc[(long long)(r0 + r) * x.ldx + x.d + (cc - x.z)];
            else if (cc == x.z + x.c_dim) val = 1.f;
          }
          v[j] = val;
        }
        put_planes(cb, g, r, v);
      }
      // ROI targets, lane-major fp32 (training only)
      if (!x.xlm) continue;
      float* xb = x.xlm + (long long)b * x.quads * 512;
      for (int u = threadIdx.x; u < 128 * x.quads; u += 256) {
        const int r = u & 127, qd = u >> 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) {
          const float* src = x.xc + (long long)(r0 + r) * x.ldx + 4 * qd;
          const int nv = x.d - 4 * qd;
          v.x = src[0]; if (nv > 1) v.y = src[1]; if (nv > 2) v.z = src[2]; if (nv > 3) v.w = src[3];
        }
        *reinterpret_cast<float4*>(xb + ((long long)qd * 128 + r) * 4) = v;
      }
    }
  }
}

}  // namespace tcp

cudaError_t set_tcp_trace(unsigned long long* buf, int step) {
  cudaError_t e = cudaMemcpyToSymbol(tcp::g_trace, &buf, sizeof(buf));
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbol(tcp::g_trace_step, &step, sizeof(step));
}

cudaError_t configure_tcp() {
  if (const char* u = getenv("NMB_TCP_SPIN_BUDGET")) {
    const long long v = atoll(u);
    cudaError_t e = cudaMemcpyToSymbol(tcp::g_spin_budget, &v, sizeof(v));
    if (e != cudaSuccess) return e;
  }
  // Forward progress of chunked work items relies on every CTA of the grid being co-resident (a chunk spins on its
  // predecessor, which may sit on another SM): the grid is capped at the SM count and needs one CTA per SM to fit.
  int per_sm = 0;
  cudaError_t e = cudaFuncSetAttribute(tcp::train_tcp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tcp::train_tcp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::kSmemBytes);
  if (e != cudaSuccess) return e;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tcp::train_tcp_kernel<false>, tcp::kThreadsP, tcp::kSmemBytes);
  if (e != cudaSuccess) return e;
  return per_sm >= 1 ? cudaSuccess : cudaErrorLaunchOutOfResources;
}

cudaError_t launch_xprep(const void* items_dev, int n_items, int max_blocks, cudaStream_t st) {
  if (n_items <= 0) return cudaSuccess;
  dim3 grid(max_blocks < 1 ? 1 : (max_blocks > 64 ? 64 : max_blocks), n_items > 1024 ? 1024 : n_items);
  tcp::xprep_kernel<<<grid, 256, 0, st>>>(static_cast<const tcp::XPrepItem*>(items_dev), n_items);
  return cudaGetLastError();
}

namespace tcp {
struct PrepArgs {
  MemberDev* members; const ProgramDev* progs; const MemberTc* mtc; float* master; long long master_floats;
  int n_members; int adam;
};
// Before the persistent kernel: BF16 planes of every member's weights (grid.y == 0) and the caller's row-major
// parameters / Adam moments -> lane-major master state.  After it (gather == 0): master state -> caller's layout.
// One CTA per (member, layer); 32-row x 128-column tiles go through shared memory so that BOTH sides are read and
// written in contiguous 128-byte (row-major side) / 512-byte (lane-major side) warp accesses.
//   kind 0 (lane = output row o):  master float4 (quad qd, lane o) = W[o][4 qd .. 4 qd + 3]
//   kind 1 (lane = input index i): master float4 (quad qd, lane i) = W[4 qd .. 4 qd + 3][i]
__global__ void __launch_bounds__(256) tcp_move_kernel(const PrepArgs a, const int gather) {
  __shared__ float T[32][129];
  const int mi = blockIdx.x, layer = blockIdx.y;
  MemberDev& mb = a.members[mi];
  const ProgramDev& pg = a.progs[mb.arch_idx];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (gather && layer == 0 && tid == 0) mb.launch_base = mb.steps_done;
  if (layer >= pg.n_mlayers) return;
  const MLayer ml = pg.mlayers[layer];
  if (!a.adam) {          // forward / backward only: just this layer's BF16 planes
    if (gather) build_weight_planes(pg, mb, a.mtc[mi], tid, blockDim.x, ml.p_off);
    return;
  }
  float* ext[3] = {mb.params, mb.adam_m, mb.adam_v};
  float* mst0 = a.master + (long long)mi * 3 * a.master_floats;
  // float4 access to the caller's rows: every row starts on a 16-byte boundary (row stride and matrix offsets are
  // multiples of 4 floats) when the three base pointers do, and a 128-column block never reads past the row stride
  const bool vec = ((reinterpret_cast<unsigned long long>(ext[0]) | reinterpret_cast<unsigned long long>(ext[1]) |
                     reinterpret_cast<unsigned long long>(ext[2])) & 15ull) == 0ull && (ml.p_ld & 3) == 0 && (ml.p_off & 3) == 0;
  for (int r0 = 0; r0 < ml.rows; r0 += 32) {
    const int nr = min(32, ml.rows - r0);
    for (int c0 = 0; c0 < ml.cols; c0 += 128) {
      const int nc = min(128, ml.cols - c0);
      for (int k = 0; k < 3; ++k) {
        float* E = ext[k] + ml.p_off + (long long)r0 * ml.p_ld + c0;
        float* M = mst0 + k * a.master_floats + ml.mst_off;
        if (gather) {
          if (vec) {            // one 512-byte warp access per row: 32 lanes x float4 (the row padding is readable)
            for (int r = warp; r < 32; r += 8) {
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (r < nr && 4 * lane < nc) v = __ldcs(reinterpret_cast<const float4*>(E + (long long)r * ml.p_ld + 4 * lane));
              const int cc = 4 * lane;
              T[r][cc] = cc < nc ? v.x : 0.f; T[r][cc + 1] = cc + 1 < nc ? v.y : 0.f;
              T[r][cc + 2] = cc + 2 < nc ? v.z : 0.f; T[r][cc + 3] = cc + 3 < nc ? v.w : 0.f;
            }
          } else {
            for (int r = warp; r < 32; r += 8)
              for (int cc = lane; cc < 128; cc += 32) T[r][cc] = (r < nr && cc < nc) ? E[(long long)r * ml.p_ld + cc] : 0.f;
          }
        } else if (ml.kind == 0) {
          // lane = row: quads of this column block, 512 contiguous bytes per warp
          for (int q = warp; q * 4 < nc; q += 8) {
            const float4 v = lane < nr ? *reinterpret_cast<const float4*>(M + ((long long)((c0 >> 2) + q) * ml.R + r0 + lane) * 4)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            T[lane][4 * q] = v.x; T[lane][4 * q + 1] = v.y; T[lane][4 * q + 2] = v.z; T[lane][4 * q + 3] = v.w;
          }
        } else {
          // lane = column: the 8 row quads of this row block
          for (int q = warp; q * 4 < nr; q += 8)
            for (int cc = lane; cc < nc; cc += 32) {
              const float4 v = *reinterpret_cast<const float4*>(M + ((long long)((r0 >> 2) + q) * ml.R + c0 + cc) * 4);
              T[4 * q][cc] = v.x; T[4 * q + 1][cc] = v.y; T[4 * q + 2][cc] = v.z; T[4 * q + 3][cc] = v.w;
            }
        }
        __syncthreads();
        if (gather && k == 0) {
          // BF16 hi/lo planes of these 32 rows x 128 columns straight from the tile (parameters are read once):
          // thread = (row, 8-column group); consecutive rows -> 512 contiguous bytes per warp store
          for (int b = 0; b < pg.n_wblocks; ++b) {
            const WBlock wb = pg.wblocks[b];
            if (wb.p_off != ml.p_off) continue;
            unsigned char* dst = a.mtc[mi].wplanes + wb.wp_off;
            if (wb.transposed) {      // thread = (input index, group of 8 output rows): 512 contiguous bytes per warp store
              for (int u = tid; u < 128 * 4; u += 256) {
                const int il = u & 127, ogl = u >> 7, i = c0 + il, og = (r0 >> 3) + ogl;
                if (i >= wb.R || og >= wb.cg) continue;
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = (8 * ogl + j < nr && il < nc) ? T[8 * ogl + j][il] : 0.f;
                uint4 h, l;
                tc::split8(x, h, l);
                unsigned char* p = dst + (long long)og * 32 * wb.R + i * 16;
                *reinterpret_cast<uint4*>(p) = h;
                *reinterpret_cast<uint4*>(p + 16 * wb.R) = l;
              }
              continue;
            }
            if (r0 < wb.row0 || r0 >= wb.row0 + wb.R) continue;
            for (int u = tid; u < 32 * 16; u += 256) {
              const int r = u & 31, gl = u >> 5, g = (c0 >> 3) + gl, rb = r0 - wb.row0 + r;
              if (g >= wb.cg || rb >= wb.R) continue;
              float x[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) x[j] = (r < nr && rb < wb.rows_valid) ? T[r][8 * gl + j] : 0.f;
              uint4 h, l;
              tc::split8(x, h, l);
              unsigned char* p = dst + (long long)g * 32 * wb.R + rb * 16;
              *reinterpret_cast<uint4*>(p) = h;
              *reinterpret_cast<uint4*>(p + 16 * wb.R) = l;
            }
          }
        }
        if (!gather) {
          if (vec) {            // whole quads; a quad that straddles the last column keeps the caller's padding (zeros)
            for (int r = warp; r < nr; r += 8) {
              const int cc = 4 * lane;
              if (cc < nc) {
                float4 v = make_float4(T[r][cc], T[r][cc + 1], T[r][cc + 2], T[r][cc + 3]);
                if (cc + 4 > nc) {
                  const float4 old = *reinterpret_cast<const float4*>(E + (long long)r * ml.p_ld + cc);
                  if (cc + 1 >= nc) v.y = old.y;
                  if (cc + 2 >= nc) v.z = old.z;
                  v.w = old.w;
                }
                __stcs(reinterpret_cast<float4*>(E + (long long)r * ml.p_ld + cc), v);
              }
            }
          } else {
            for (int r = warp; r < nr; r += 8)
              for (int cc = lane; cc < nc; cc += 32) E[(long long)r * ml.p_ld + cc] = T[r][cc];
          }
        } else if (ml.kind == 0) {
          for (int q = warp; q * 4 < nc; q += 8)
            if (lane < nr)
              *reinterpret_cast<float4*>(M + ((long long)((c0 >> 2) + q) * ml.R + r0 + lane) * 4) =
                  make_float4(T[lane][4 * q], T[lane][4 * q + 1], T[lane][4 * q + 2], T[lane][4 * q + 3]);
        } else {
          for (int q = warp; q * 4 < nr; q += 8)
            for (int cc = lane; cc < nc; cc += 32)
              *reinterpret_cast<float4*>(M + ((long long)((r0 >> 2) + q) * ml.R + c0 + cc) * 4) =
                  make_float4(T[4 * q][cc], T[4 * q + 1][cc], T[4 * q + 2][cc], T[4 * q + 3][cc]);
        }
        __syncthreads();
      }
    }
  }
}
}  // namespace tcp

// Forward-only (reconstruction) launch of the pipelined kernel: `progs` / tables are those of the forward-only programs.
cudaError_t launch_recon_tcp(const TrainLaunch& t, const tcp::ProgramDev* progs, const tcp::ProgramDev* train_progs,
                             const tcp::MemberTc* mtc, unsigned char* stash, long long stash_bytes,
                             const tcp::MStep* msteps, const int* ms_off, const int* ms_cnt, int n_archs,
                             const tcp::EpiP* epis_p, const int* ep_off, const int* ep_cnt, const unsigned char* ep_first,
                             int max_mlayers, int recon_mode, const tcp::ReconTc* rtc, const tcp::ReconWork* rwork, int n_rwork,
                             int n_sm, cudaStream_t st) {
  if (n_rwork <= 0) return cudaSuccess;
  const int grid = n_rwork < n_sm ? n_rwork : n_sm;
  if (grid < n_rwork) {
    cudaError_t e = cudaMemsetAsync(t.work_counter, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
  }
  // the caller may have changed the parameters since the last call: rebuild the BF16 planes (layout of the TRAINING
  // program, shared by both programs) from the fp32 parameters (max_mlayers == 0: the caller vouches for the planes)
  if (max_mlayers > 0) {
    tcp::PrepArgs pa{t.members, train_progs, mtc, nullptr, 0, t.n_members, 0};
    const dim3 pgrid((unsigned)t.n_members, (unsigned)max_mlayers);
    tcp::tcp_move_kernel<<<pgrid, 256, 0, st>>>(pa, 1);
  }
  tcp::LaunchP L;
  L.n_chunks = 1;
  L.t = t; L.progs = progs; L.mtc = mtc; L.stash = stash; L.stash_bytes = stash_bytes;
  L.master = nullptr; L.master_floats = 0;
  L.recon_mode = recon_mode; L.n_rwork = n_rwork; L.rtc = rtc; L.rwork = rwork;
  for (int a = 0; a < n_archs; ++a) {
    L.ms_off[a] = ms_off[a]; L.ms_cnt[a] = ms_cnt[a]; L.ep_off[a] = ep_off[a]; L.ep_cnt[a] = ep_cnt[a];
    for (int g = 0; g < 4; ++g) L.ep_first[a][g] = ep_first[4 * a + g];
  }
  { const char* u = getenv("NMB_TCP_JUMP"); L.jump = !(u && u[0] == '0'); }
  {
    int n_ep = 0;
    for (int a = 0; a < n_archs; ++a) if (ep_cnt[a] > 0 && ep_off[a] + ep_cnt[a] > n_ep) n_ep = ep_off[a] + ep_cnt[a];
    if (n_ep > 0) std::memcpy(L.epis_p, epis_p, sizeof(tcp::EpiP) * (size_t)n_ep);
  }
  std::memcpy(L.msteps, msteps, sizeof(tcp::MStep) * (size_t)(ms_off[n_archs - 1] + ms_cnt[n_archs - 1]));
  tcp::train_tcp_kernel<true><<<grid, tcp::kThreadsP, tcp::kSmemBytes, st>>>(L);
  return cudaGetLastError();
}

namespace tcp {
// resident state: the only per-launch bookkeeping of the state-conversion pass that is still needed
__global__ void tcp_launch_base_kernel(MemberDev* members, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) members[i].launch_base = members[i].steps_done;
}
}  // namespace tcp

// master state (lane-major p, m, v of every member) -> the caller's row-major buffers
cudaError_t launch_tcp_scatter(MemberDev* members, int n_members, const tcp::ProgramDev* progs, const tcp::MemberTc* mtc,
                               float* master, long long master_floats, int max_mlayers, cudaStream_t st) {
  tcp::PrepArgs pa{members, progs, mtc, master, master_floats, n_members, 1};
  const dim3 pgrid((unsigned)n_members, (unsigned)max_mlayers);
  tcp::tcp_move_kernel<<<pgrid, 256, 0, st>>>(pa, 0);
  return cudaGetLastError();
}

// gather_in: convert the caller's buffers to the master layout (and build the weight planes) before the launch;
// scatter_out: convert back after it.  Both false = the state stays resident between consecutive calls.
cudaError_t launch_train_tcp(const TrainLaunch& t, const tcp::ProgramDev* progs, const tcp::MemberTc* mtc,
                             unsigned char* stash, long long stash_bytes, float* master, long long master_floats,
                             const tcp::MStep* msteps, const int* ms_off, const int* ms_cnt, int n_archs,
                             const tcp::EpiP* epis_p, const int* ep_off, const int* ep_cnt, const unsigned char* ep_first,
                             int max_mlayers, int n_sm, bool gather_in, bool scatter_out, cudaStream_t st) {
  if (t.n_members <= 0 || t.n_steps <= 0) return cudaSuccess;
  // more members than SMs: deal chunks of >= 4 steps, so that the launch does not end on a few whole members
  int n_chunks = 1;
  if (t.n_members > n_sm) {
    const long long ns_min = t.n_is_epochs ? t.n_steps : t.n_steps;      // epochs mode: every member has >= n_steps steps
    n_chunks = (int)(ns_min / 4 < 8 ? ns_min / 4 : 8);
    if (n_chunks < 1) n_chunks = 1;
    if (const char* u = getenv("NMB_TCP_CHUNKS")) { const int v = atoi(u); if (v >= 1 && v <= t.n_steps) n_chunks = v; }   // dev knob
  }
  const long long n_items = (long long)t.n_members * n_chunks;
  const int grid = n_items < n_sm ? (int)n_items : n_sm;
  if (grid < n_items) {
    cudaError_t e = cudaMemsetAsync(t.work_counter, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
  }
  tcp::PrepArgs pa{t.members, progs, mtc, master, master_floats, t.n_members, (t.flags & NMB_TRAIN_NO_ADAM) ? 0 : 1};
  const dim3 pgrid((unsigned)t.n_members, (unsigned)max_mlayers);
  if (gather_in) tcp::tcp_move_kernel<<<pgrid, 256, 0, st>>>(pa, 1);
  else tcp::tcp_launch_base_kernel<<<(t.n_members + 255) / 256, 256, 0, st>>>(t.members, t.n_members);
  tcp::LaunchP L;
  L.n_chunks = n_chunks;
  L.t = t; L.progs = progs; L.mtc = mtc; L.stash = stash; L.stash_bytes = stash_bytes;
  L.master = master; L.master_floats = master_floats;
  L.recon_mode = 0; L.n_rwork = 0; L.rtc = nullptr; L.rwork = nullptr;
  for (int a = 0; a < n_archs; ++a) {
    L.ms_off[a] = ms_off[a]; L.ms_cnt[a] = ms_cnt[a]; L.ep_off[a] = ep_off[a]; L.ep_cnt[a] = ep_cnt[a];
    for (int g = 0; g < 4; ++g) L.ep_first[a][g] = ep_first[4 * a + g];
  }
  { const char* u = getenv("NMB_TCP_JUMP"); L.jump = !(u && u[0] == '0'); }
  {
    int n_ep = 0;
    for (int a = 0; a < n_archs; ++a) if (ep_cnt[a] > 0 && ep_off[a] + ep_cnt[a] > n_ep) n_ep = ep_off[a] + ep_cnt[a];
    if (n_ep > 0) std::memcpy(L.epis_p, epis_p, sizeof(tcp::EpiP) * (size_t)n_ep);
  }
  std::memcpy(L.msteps, msteps, sizeof(tcp::MStep) * (size_t)(ms_off[n_archs - 1] + ms_cnt[n_archs - 1]));
  tcp::train_tcp_kernel<false><<<grid, tcp::kThreadsP, tcp::kSmemBytes, st>>>(L);
  if (pa.adam && scatter_out) tcp::tcp_move_kernel<<<pgrid, 256, 0, st>>>(pa, 0);
  return cudaGetLastError();
}

}  // namespace nmb
