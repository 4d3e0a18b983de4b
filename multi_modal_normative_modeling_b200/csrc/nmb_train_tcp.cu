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
#include "nmb_tc_gemm.cuh"
#include "nmb_internal.h"
#include "nmb_tcp.h"
#include "nmb_fusion.cuh"

namespace nmb {
namespace tcp {

struct Ctrl {
  uint64_t full[kSlots], empty[kSlots], accbar[4];
  uint32_t tmem;
  volatile uint32_t epi_done;
  int member;
  float red[40];
};

__device__ __forceinline__ uint32_t ld_acquire(const volatile uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(tc::smem_u32((const void*)p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(volatile uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(tc::smem_u32((const void*)p)), "r"(v) : "memory");
}
// Bounded spin (a protocol bug must trap, not hang the GPU).
__device__ __forceinline__ void wait_epi(const volatile uint32_t* p, uint32_t need) {
  if (ld_acquire(p) >= need) return;
  const long long t0 = clock64();
  while (ld_acquire(p) < need) {
    __nanosleep(32);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

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
__device__ __forceinline__ StepVars step_vars(const MemberDev& mb, long long s, long long i, int n_epis) {
  StepVars v;
  const int spe = (mb.n_rows + mb.batch - 1) / mb.batch;
  v.pos = (int)(s % spe);
  v.row0 = v.pos * mb.batch;
  v.rows = min(mb.batch, mb.n_rows - v.row0);
  v.rows_h[0] = min(v.rows, 128);
  v.rows_h[1] = v.rows - v.rows_h[0];
  v.base = 1u + (uint32_t)i * (uint32_t)n_epis;
  return v;
}

// ------------------------------------------------------------------------------------------------
// producer: TMA bulk copies of operand tiles, in program order, into the ring
__device__ void producer_role(const TrainLaunch& t, const ProgramDev& pg, const MemberDev& mb, const MemberTc& mt,
                              const unsigned char* stash, unsigned char* smem, Ctrl* ctl, uint32_t& seq) {
  unsigned char* ring = smem + kSmemRing;
  for (long long i = 0; i < t.n_steps; ++i) {
    const StepVars sv = step_vars(mb, mb.steps_done + i, i, pg.n_epis);
    for (int k = 0; k < pg.n_steps; ++k) {
      const Step& st = pg.steps[k];
      if (sv.rows_h[st.half] == 0) continue;
      uint32_t need = st.dep ? sv.base + (uint32_t)st.dep : 0u;
      if (st.b_space == SP_W) need = max(need, sv.base);
      if (need) wait_epi(&ctl->epi_done, need);
      for (int which = 0; which < 2; ++which) {
        const int space = which == 0 ? st.a_space : st.b_space;
        const uint32_t bytes = which == 0 ? st.a_bytes : st.b_bytes;
        if (which == 0 && bytes == 0) continue;
        const long long off = which == 0 ? st.a_off : st.b_off;
        const unsigned char* src;
        if (space == SP_W) src = mt.wplanes + off;
        else if (space == SP_STASH) src = stash + off;
        else src = mt.xplanes[st.x_mod] + ((long long)(sv.pos * mt.n_half + st.half) * pg.lay.x_cg[st.x_mod]) * 4096 + off;
        const uint32_t slot = seq % kSlots, use = seq / kSlots;
        tc::mbar_wait(&ctl->empty[slot], (use & 1u) ^ 1u);
        expect_tx(&ctl->full[slot], bytes);
        bulk_load(ring + slot * kSlotBytes, src, bytes, &ctl->full[slot]);
        ++seq;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// MMA issuer: 3 BF16 passes per K = 16 step
__device__ void mma_role(const TrainLaunch& t, const ProgramDev& pg, const MemberDev& mb, unsigned char* smem,
                         Ctrl* ctl, uint32_t tmem, uint32_t& seq) {
  const uint32_t ring = tc::smem_u32(smem + kSmemRing);
  const uint32_t act0 = tc::smem_u32(smem);
  for (long long i = 0; i < t.n_steps; ++i) {
    const StepVars sv = step_vars(mb, mb.steps_done + i, i, pg.n_epis);
    wait_epi(&ctl->epi_done, sv.base);
    for (int k = 0; k < pg.n_steps; ++k) {
      const Step& st = pg.steps[k];
      if (sv.rows_h[st.half] == 0) continue;
      if (st.mma_dep) wait_epi(&ctl->epi_done, sv.base + (uint32_t)st.mma_dep);
      uint32_t a_base, slot_a = 0xFFFFFFFFu;
      if (st.a_bytes) {
        slot_a = seq % kSlots;
        tc::mbar_wait(&ctl->full[slot_a], (seq / kSlots) & 1u);
        a_base = ring + slot_a * kSlotBytes;
        ++seq;
      } else {
        a_base = act0 + st.half * kActBytes + st.a_start;
      }
      const uint32_t slot_b = seq % kSlots;
      tc::mbar_wait(&ctl->full[slot_b], (seq / kSlots) & 1u);
      const uint32_t b_base = ring + slot_b * kSlotBytes;
      ++seq;
      tc::fence_after();
      const uint32_t idesc = tc::make_idesc(st.n, st.a_mn, st.b_mn);
      const uint32_t d = tmem + st.tmem_col;
      for (int ks = 0; ks < st.ksteps; ++ks) {
        const uint32_t a_hi = a_base + ks * st.a_kadv, b_hi = b_base + ks * st.b_kadv;
        const uint64_t da_hi = make_desc(a_hi, st.a_lbo, st.a_sbo), da_lo = make_desc(a_hi + st.a_lo, st.a_lbo, st.a_sbo);
        const uint64_t db_hi = make_desc(b_hi, st.b_lbo, st.b_sbo), db_lo = make_desc(b_hi + st.b_lo, st.b_lbo, st.b_sbo);
        tc::mma_bf16(d, da_hi, db_hi, idesc, (st.first && ks == 0) ? 0u : 1u);
        tc::mma_bf16(d, da_lo, db_hi, idesc, 1u);
        tc::mma_bf16(d, da_hi, db_lo, idesc, 1u);
      }
      tc::mma_commit(&ctl->empty[slot_b]);
      if (slot_a != 0xFFFFFFFFu) tc::mma_commit(&ctl->empty[slot_a]);
      if (st.commit == 1 || (st.commit == 2 && sv.rows_h[1] == 0)) tc::mma_commit(&ctl->accbar[st.commit_buf]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// epilogue helpers
struct EpiCtx {
  const ArchDesc* a; const ProgramDev* pg; MemberDev* mb; const MemberTc* mt;
  unsigned char* smem; unsigned char* stash; float* scratch; Ctrl* ctl;
  uint32_t tmem;
  int warp, lane, row, cpart, tid;
  unsigned flags;
  StepVars sv;
  long long step;           // global 0-based minibatch step of this member (Philox counter, Adam t - 1)
  float step_size, bc2_sqrt, b1, b2, aeps;
  float kl_acc, ll_acc;
  float dw_acc[NMB_MAX_MOD];
};

// 8 values -> hi / lo planes at (group g, row) of a 128-row block
__device__ __forceinline__ void put_planes(unsigned char* blk, int g, int row, const float (&x)[8]) {
  uint4 h, l;
  tc::split8(x, h, l);
  unsigned char* p = blk + (long long)g * 4096 + row * 16;
  *reinterpret_cast<uint4*>(p) = h;
  *reinterpret_cast<uint4*>(p + 2048) = l;
}

__device__ __forceinline__ float adam_update(const EpiCtx& c, float& m1, float& v1, float p0, float g) {
  m1 = c.b1 * m1 + (1.f - c.b1) * g;
  v1 = c.b2 * v1 + (1.f - c.b2) * g * g;
  return p0 - c.step_size * (m1 / (sqrtf(v1) / c.bc2_sqrt + c.aeps));
}
__device__ __forceinline__ void adam_scalar(const EpiCtx& c, long long idx, float g) {
  MemberDev& mb = *c.mb;
  if (c.flags & NMB_TRAIN_WRITE_GRADS) mb.grads[idx] = g;
  if (!(c.flags & NMB_TRAIN_NO_ADAM)) {
    float m1 = mb.adam_m[idx], v1 = mb.adam_v[idx];
    const float p1 = adam_update(c, m1, v1, mb.params[idx], g);
    mb.adam_m[idx] = m1; mb.adam_v[idx] = v1; mb.params[idx] = p1;
  }
}

__device__ __forceinline__ uint32_t taddr(const EpiCtx& c, int col) {
  return c.tmem + ((uint32_t)((c.warp & 3) << 5) << 16) + (uint32_t)col;
}

__device__ __forceinline__ float block_sum_epi(EpiCtx& c, float v) {
  v = warp_sum(v);
  epi_bar();
  if (c.lane == 0) c.ctl->red[c.warp] = v;
  epi_bar();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kEpiWarps; ++i) s += c.ctl->red[i];
  return s;
}

// fp32 parameters -> BF16 hi/lo planes (member start)
__device__ void build_weight_planes(const EpiCtx& c) {
  const ProgramDev& pg = *c.pg;
  const float* P = c.mb->params;
  for (int b = 0; b < pg.n_wblocks; ++b) {
    const WBlock wb = pg.wblocks[b];
    const int units = wb.R * wb.cg;
    unsigned char* dst = c.mt->wplanes + wb.wp_off;
    for (int u = c.tid; u < units; u += kEpiThreads) {
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

__device__ void epi_hidden(EpiCtx& c, const Epi& e) {
  const int h = e.half;
  const bool vr = c.row < c.sv.rows_h[h];
  unsigned char* act = c.smem + h * kActBytes;
  unsigned char* st = c.stash + e.stash_off;
  const int nl = c.a->non_linear;
  for (int ch = c.cpart; ch * 16 < e.n_cols; ch += 2) {
    const int col = ch * 16;
    float v[16];
    if (col < e.n_mma) tc::tmem_ld16(taddr(c, e.tmem_col + col), v);
    else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int cc = col + j;
      float x = v[j];
      x = (nl && x <= 0.f) ? kSlope * x : x;
      v[j] = !vr ? 0.f : (cc < e.n_valid ? x : (cc == e.n_valid ? 1.f : 0.f));
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = v[8 * q + j];
      put_planes(act, 2 * ch + q, c.row, x);
      put_planes(st, 2 * ch + q, c.row, x);
    }
  }
}

__device__ void epi_head(EpiCtx& c, const Epi& e) {
  const int h = e.half;
  const bool vr = c.row < c.sv.rows_h[h];
  const int ld = c.pg->lay.ld_mulv;
  float* dst = reinterpret_cast<float*>(c.stash + c.pg->lay.mulv[e.mod]) + (long long)(128 * h + c.row) * ld;
  for (int ch = c.cpart; ch * 16 < e.n_cols; ch += 2) {
    float v[16];
    __syncwarp();
    tc::tmem_ld16(taddr(c, e.tmem_col + ch * 16), v);
    if (vr) {
#pragma unroll
      for (int j = 0; j < 16; ++j) if (ch * 16 + j < e.n_valid) dst[ch * 16 + j] = v[j];
    }
  }
}

// fusion + reparameterisation + KL for half h; builds decoder inputs [z | c | 1] (cVAE.py:1130-1164, 199)
__device__ void epi_latent(EpiCtx& c, const Epi& e, const float* eps_src) {
  const ArchDesc& a = *c.a;
  const Layout& lay = c.pg->lay;
  const int h = e.half, Z = a.Z, M = a.M, rows = c.sv.rows_h[h];
  float* S = c.scratch;
  float* zbuf = reinterpret_cast<float*>(c.stash + lay.zbuf);
  float w[NMB_MAX_MOD];
  if (M > 1 && a.combine == NMB_COMBINE_GPOE) softmax_alpha(c.mb->params + a.alpha_off, M, w);
  const int n = rows * Z;
  const int g_base = (128 * h * Z) / 4;
  for (int g = c.tid; g * 4 < n; g += kEpiThreads) {
    float nrm[4] = {0.f, 0.f, 0.f, 0.f};
    if (!eps_src) philox_normal4(c.mb->seed, (unsigned long long)c.step, 0u, (uint32_t)(g_base + g), nrm);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int el = g * 4 + j;
      if (el >= n) break;
      const int b = el / Z, z = el - b * Z;
      const int gb = 128 * h + b;
      float mu[NMB_MAX_MOD], lv[NMB_MAX_MOD];
      for (int m = 0; m < M; ++m) {
        const float* hd = reinterpret_cast<const float*>(c.stash + lay.mulv[m]) + (long long)gb * lay.ld_mulv;
        mu[m] = hd[z]; lv[m] = hd[Z + z];
      }
      const Fused f = fuse_forward(mu, lv, M, a.combine, w);
      const float eps = eps_src ? eps_src[gb * Z + z] : nrm[j];
      S[a.s_mub + gb * Z + z] = f.mu; S[a.s_lvb + gb * Z + z] = f.lv; S[a.s_eps + gb * Z + z] = eps;
      zbuf[gb * Z + z] = f.mu + eps * expf(0.5f * f.lv);
      c.kl_acc += -0.5f * (1.f + f.lv - f.mu * f.mu - expf(f.lv));
    }
  }
  epi_bar();
  const int cg = round16(Z + a.C + 1) / 8;
  for (int m = 0; m < M; ++m) {
    const ModDesc& q = a.mod[m];
    unsigned char* st = c.stash + lay.g0[m][h];
    const float* xc = c.mb->xc[m] + (long long)(c.sv.row0 + 128 * h) * q.ldx + q.D;
    for (int u = c.tid; u < 128 * cg; u += kEpiThreads) {
      const int r = u & 127, g = u >> 7;
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cc = 8 * g + j;
        float val = 0.f;
        if (r < rows) {
          if (cc < Z) val = zbuf[(128 * h + r) * Z + cc];
          else if (cc < Z + a.C) val = xc[(long long)r * q.ldx + (cc - Z)];
          else if (cc == Z + a.C) val = 1.f;
        }
        x[j] = val;
      }
      put_planes(st, g, r, x);
      if (m == 0) put_planes(c.smem + h * kActBytes, g, r, x);
    }
  }
}

__device__ void epi_copy(EpiCtx& c, const Epi& e) {
  const uint4* src = reinterpret_cast<const uint4*>(c.stash + e.src_off);
  uint4* dst = reinterpret_cast<uint4*>(c.smem + e.half * kActBytes);
  const int n = e.src_cg * 256;       // 16-byte units
  for (int u = c.tid; u < n; u += kEpiThreads) dst[u] = src[u];
}

__device__ void epi_recon(EpiCtx& c, const Epi& e) {
  const ArchDesc& a = *c.a;
  const ModDesc& q = a.mod[e.mod];
  const Layout& lay = c.pg->lay;
  const int h = e.half;
  const bool vr = c.row < c.sv.rows_h[h];
  const int gauss = a.loss_kind == NMB_LOSS_GAUSS_LL;
  const float inv_rows = 1.f / c.sv.rows;
  const float inv_rows_d = inv_rows / q.D;
  const float ll_scale = gauss ? inv_rows : inv_rows_d;
  const float* P = c.mb->params;
  unsigned char* st = c.stash + e.stash_off;
  const int grow = c.sv.row0 + 128 * h + c.row;
  const float* xrow = c.mb->xc[e.mod] + (long long)grow * q.ldx;
  float* keep = (c.flags & NMB_TRAIN_KEEP_ACTS) ? c.scratch + q.s_xr + (long long)(128 * h + c.row) * q.ld_xh : nullptr;
  float* lampart = reinterpret_cast<float*>(c.stash + lay.lampart[e.mod]) + (long long)(h * 4 + (c.warp & 3)) * round4(q.D);
  for (int ch = c.cpart; ch < 4; ch += 2) {
    const int col = ch * 16, gc = e.col0 + col;
    int nv = e.n_valid - col; nv = nv < 0 ? 0 : (nv > 16 ? 16 : nv);
    float v[16], xt[16], l[16], gr[16], qv[16];
    __syncwarp();
    tc::tmem_ld16(taddr(c, e.tmem_col + col), v);
#pragma unroll
    for (int j = 0; j < 16; ++j) { xt[j] = 0.f; l[j] = 0.f; }
    if (vr && nv > 0) load16(xrow + gc, nv, 0.f, xt);
    if (gauss && nv > 0) load16(P + q.lam_off + gc, nv, 0.f, l);
    float ll = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const bool ok = vr && j < nv;
      const float r = xt[j] - v[j];
      float t, g, qq = 0.f;
      if (gauss) {
        const float iv = __expf(-l[j]);
        t = -0.5f * r * r * iv - 0.5f * l[j] - 0.5f * kLog2Pi;
        g = -r * iv * inv_rows;
        qq = 0.5f * (1.f - r * r * iv);
      } else {
        t = -r * r;
        g = -2.f * r * inv_rows_d;
      }
      ll += ok ? t : 0.f;
      gr[j] = ok ? g : 0.f;
      qv[j] = ok ? qq : 0.f;
    }
    c.ll_acc += ll * ll_scale;
#pragma unroll
    for (int qd = 0; qd < 2; ++qd) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = gr[8 * qd + j];
      put_planes(st, 2 * ch + qd, c.row, x);
    }
    if (keep && vr && nv > 0) store16(keep + gc, nv, v);
    if (gauss) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float s = warp_sum(qv[j]);
        if (c.lane == j && j < nv) lampart[gc + j] = s;
      }
    }
  }
}

__device__ void epi_lam(EpiCtx& c, const Epi& e) {
  const ModDesc& q = c.a->mod[e.mod];
  const float* part = reinterpret_cast<const float*>(c.stash + c.pg->lay.lampart[e.mod]);
  const int ld = round4(q.D);
  const int np = c.sv.rows_h[1] > 0 ? 8 : 4;
  const float inv_rows = 1.f / c.sv.rows;
  for (int n = c.tid; n < q.D; n += kEpiThreads) {
    float s = 0.f;
    for (int k = 0; k < np; ++k) s += part[k * ld + n];
    adam_scalar(c, q.lam_off + n, s * inv_rows);
  }
}

__device__ void epi_dgrad(EpiCtx& c, const Epi& e) {
  const int h = e.half;
  const bool vr = c.row < c.sv.rows_h[h];
  unsigned char* act = c.smem + h * kActBytes;
  const unsigned char* sg = c.stash + e.src_off;
  const int nl = c.a->non_linear;
  for (int ch = c.cpart; ch * 16 < e.n_cols; ch += 2) {
    const int col = ch * 16;
    uint4 s0 = make_uint4(0, 0, 0, 0), s1 = s0;
    if (nl) {
      s0 = *reinterpret_cast<const uint4*>(sg + (long long)(2 * ch) * 4096 + c.row * 16);
      s1 = *reinterpret_cast<const uint4*>(sg + (long long)(2 * ch + 1) * 4096 + c.row * 16);
    }
    float v[16];
    if (col < e.n_mma) tc::tmem_ld16(taddr(c, e.tmem_col + col), v);
    else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.f;
    }
    const uint32_t sw[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t hb = (sw[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;      // bf16 bits of the stored activation
      const bool nonpos = nl && ((hb & 0x8000u) || (hb & 0x7FFFu) == 0u);
      const float x = nonpos ? kSlope * v[j] : v[j];
      v[j] = (vr && col + j < e.n_valid) ? x : 0.f;
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

__device__ void epi_dz(EpiCtx& c, const Epi& e) {
  const int h = e.half, Z = c.a->Z;
  const bool vr = c.row < c.sv.rows_h[h];
  float* dz = reinterpret_cast<float*>(c.stash + c.pg->lay.dz) + (long long)(128 * h + c.row) * Z;
  for (int ch = c.cpart; ch * 16 < e.n_mma; ch += 2) {
    float v[16];
    __syncwarp();
    tc::tmem_ld16(taddr(c, e.tmem_col + ch * 16), v);
    if (vr) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int cc = ch * 16 + j;
        if (cc < Z) dz[cc] = e.mod > 0 ? dz[cc] + v[j] : v[j];
      }
    }
  }
}

// latent + fusion backward for half h -> d[mu | logvar] planes per modality
__device__ void epi_latent_bwd(EpiCtx& c, const Epi& e) {
  const ArchDesc& a = *c.a;
  const Layout& lay = c.pg->lay;
  const int h = e.half, Z = a.Z, M = a.M, rows = c.sv.rows_h[h];
  const float* S = c.scratch;
  const float* P = c.mb->params;
  const float* dzb = reinterpret_cast<const float*>(c.stash + lay.dz);
  const float inv_rows = 1.f / c.sv.rows;
  float w[NMB_MAX_MOD];
  const bool gpoe = M > 1 && a.combine == NMB_COMBINE_GPOE;
  if (gpoe) softmax_alpha(P + a.alpha_off, M, w);
  for (int el = c.tid; el < rows * Z; el += kEpiThreads) {
    const int b = el / Z, z = el - b * Z;
    const int gi = (128 * h + b) * Z + z;
    const float mub = S[a.s_mub + gi], lvb = S[a.s_lvb + gi], eps = S[a.s_eps + gi], dz = dzb[gi];
    const float sd = expf(0.5f * lvb);
    const float dmu_bar = dz + M * mub * inv_rows;
    const float dlv_bar = dz * eps * sd * 0.5f + M * (expf(lvb) - 1.f) * 0.5f * inv_rows;
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
  epi_bar();
  const int cg = round16(2 * Z) / 8;
  for (int m = 0; m < M; ++m) {
    const float* hd0 = reinterpret_cast<const float*>(c.stash + lay.mulv[m]) + (long long)(128 * h) * lay.ld_mulv;
    unsigned char* st = c.stash + lay.dmulv[m][h];
    for (int u = c.tid; u < 128 * cg; u += kEpiThreads) {
      const int r = u & 127, g = u >> 7;
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cc = 8 * g + j;
        x[j] = (r < rows && cc < 2 * Z) ? hd0[(long long)r * lay.ld_mulv + cc] : 0.f;
      }
      if (m == 0) put_planes(c.smem + h * kActBytes, g, r, x);
      else put_planes(st, g, r, x);
    }
  }
}

// weight gradient (lane = output row) fused with Adam; rewrites the BF16 planes of the layer
__device__ void epi_wgrad(EpiCtx& c, const Epi& e) {
  MemberDev& mb = *c.mb;
  const int o = c.row;
  const bool vo = o < e.p_rows;
  unsigned char* wp = c.mt->wplanes + e.wp_off;
  const int wp_cg = round16(e.p_cols) / 8;
  for (int ch = c.cpart; ch * 16 < e.n_mma; ch += 2) {
    const int col = e.col0 + ch * 16;
    float g[16];
    __syncwarp();
    tc::tmem_ld16(taddr(c, e.tmem_col + ch * 16), g);
    int nv = e.p_ld - col; nv = nv < 0 ? 0 : (nv > 16 ? 16 : nv);
    if (vo && nv > 0) {
      const long long base = e.p_off + (long long)o * e.p_ld + col;
      if (c.flags & NMB_TRAIN_WRITE_GRADS) store16(mb.grads + base, nv, g);
      if (!(c.flags & NMB_TRAIN_NO_ADAM)) {
        float p0[16], m1[16], v1[16];
        load16(mb.params + base, nv, 0.f, p0);
        load16(mb.adam_m + base, nv, 0.f, m1);
        load16(mb.adam_v + base, nv, 0.f, v1);
#pragma unroll
        for (int j = 0; j < 16; ++j) p0[j] = adam_update(c, m1[j], v1[j], p0[j], g[j]);
        store16(mb.adam_m + base, nv, m1);
        store16(mb.adam_v + base, nv, v1);
        store16(mb.params + base, nv, p0);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int gi = col / 8 + q;
          if (gi < wp_cg) {
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = (col + 8 * q + j < e.p_cols) ? p0[8 * q + j] : 0.f;
            uint4 hh, ll;
            tc::split8(x, hh, ll);
            unsigned char* p = wp + (long long)gi * 32 * e.wp_R + o * 16;
            *reinterpret_cast<uint4*>(p) = hh;
            *reinterpret_cast<uint4*>(p + 16 * e.wp_R) = ll;
          }
        }
      }
    }
  }
}

// transposed weight gradient of decoder_mean_layer: lane = input index i, columns = output rows o
__device__ void epi_wgrad_t(EpiCtx& c, const Epi& e) {
  MemberDev& mb = *c.mb;
  const int i = c.row;
  const bool vi = i < e.p_cols;
  unsigned char* wp = c.mt->wplanes + e.wp_off;
  const long long blk_bytes = (long long)e.src_cg * 2048;     // one 64-row planes block
  for (int ch = c.cpart; ch * 16 < e.n_mma; ch += 2) {
    float g[16];
    __syncwarp();
    tc::tmem_ld16(taddr(c, e.tmem_col + ch * 16), g);
    if (vi) {
      const int o0 = e.col0 + ch * 16;
      float p0[16], m1[16], v1[16];
      const bool adam = !(c.flags & NMB_TRAIN_NO_ADAM);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const long long idx = e.p_off + (long long)(o0 + j) * e.p_ld + i;
        const bool ok = o0 + j < e.p_rows;
        if (ok && (c.flags & NMB_TRAIN_WRITE_GRADS)) mb.grads[idx] = g[j];
        p0[j] = (ok && adam) ? mb.params[idx] : 0.f;
        m1[j] = (ok && adam) ? mb.adam_m[idx] : 0.f;
        v1[j] = (ok && adam) ? mb.adam_v[idx] : 0.f;
      }
      if (adam) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int o = o0 + j;
          if (o < e.p_rows) {
            const long long idx = e.p_off + (long long)o * e.p_ld + i;
            const float p1 = adam_update(c, m1[j], v1[j], p0[j], g[j]);
            mb.adam_m[idx] = m1[j]; mb.adam_v[idx] = v1[j]; mb.params[idx] = p1;
            const __nv_bfloat16 hb = __float2bfloat16_rn(p1);
            const __nv_bfloat16 lb = __float2bfloat16_rn(p1 - __bfloat162float(hb));
            unsigned char* p = wp + (long long)(o >> 6) * blk_bytes + (long long)(i >> 3) * 2048 + (o & 63) * 16 + (i & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(p) = hb;
            *reinterpret_cast<__nv_bfloat16*>(p + 1024) = lb;
          }
        }
      }
    }
  }
}

__device__ void epi_step_end(EpiCtx& c, float* loss_out) {
  const ArchDesc& a = *c.a;
  const int M = a.M;
  const float kl = block_sum_epi(c, c.kl_acc) / c.sv.rows;
  const float ll = block_sum_epi(c, c.ll_acc);
  if (loss_out && c.tid == 0) { loss_out[0] = M * kl - ll; loss_out[1] = M * kl; loss_out[2] = ll; }
  if (M > 1 && a.combine == NMB_COMBINE_GPOE) {
    float w[NMB_MAX_MOD], dw_tot[NMB_MAX_MOD];
    softmax_alpha(c.mb->params + a.alpha_off, M, w);
    for (int m = 0; m < M; ++m) dw_tot[m] = block_sum_epi(c, c.dw_acc[m]);
    epi_bar();
    if (c.tid == 0) {
      float dot = 0.f;
      for (int m = 0; m < M; ++m) dot += w[m] * dw_tot[m];
      for (int m = 0; m < M; ++m) adam_scalar(c, a.alpha_off + m, w[m] * (dw_tot[m] - dot));
    }
  }
  c.kl_acc = 0.f; c.ll_acc = 0.f;
  for (int m = 0; m < M; ++m) c.dw_acc[m] = 0.f;
}

__device__ void epilogue_role(const TrainLaunch& t, int mi, EpiCtx& c, uint32_t& acc_par) {
  const ProgramDev& pg = *c.pg;
  MemberDev& mb = *c.mb;
  c.kl_acc = 0.f; c.ll_acc = 0.f;
  for (int m = 0; m < NMB_MAX_MOD; ++m) c.dw_acc[m] = 0.f;
  build_weight_planes(c);
  fence_async_all();
  epi_bar();
  if (c.tid == 0) st_release(&c.ctl->epi_done, 1u);
  const long long s0 = mb.steps_done;
  for (long long i = 0; i < t.n_steps; ++i) {
    const long long s = s0 + i;
    c.sv = step_vars(mb, s, i, pg.n_epis);
    c.step = s;
    {
      const double tt = (double)(s + 1);
      const float lr = mb.lr_steps ? mb.lr_steps[s] : mb.lr;
      c.step_size = (float)((double)lr / (1.0 - pow((double)mb.beta1, tt)));
      c.bc2_sqrt = (float)sqrt(1.0 - pow((double)mb.beta2, tt));
    }
    const float* eps = t.eps_override ? t.eps_override + ((long long)mi * t.n_steps + i) * mb.batch * c.a->Z : nullptr;
    float* lo = t.loss_out ? t.loss_out + ((long long)mi * t.n_steps + i) * 3 : nullptr;
    for (int k = 0; k < pg.n_epis; ++k) {
      const Epi& e = pg.epis[k];
      const bool active = e.half == 2 ? true : c.sv.rows_h[e.half] > 0;
      if (!active) continue;
      if (e.buf >= 0) {
        tc::mbar_wait(&c.ctl->accbar[e.buf], (acc_par >> e.buf) & 1u);
        acc_par ^= 1u << e.buf;
        tc::fence_after();
      }
      switch (e.kind) {
        case EK_HIDDEN: epi_hidden(c, e); break;
        case EK_HEAD: epi_head(c, e); break;
        case EK_LATENT: epi_latent(c, e, eps); break;
        case EK_COPY: epi_copy(c, e); break;
        case EK_RECON: epi_recon(c, e); break;
        case EK_LAM: epi_lam(c, e); break;
        case EK_DGRAD: epi_dgrad(c, e); break;
        case EK_DZ: epi_dz(c, e); break;
        case EK_LATENT_BWD: epi_latent_bwd(c, e); break;
        case EK_WGRAD: epi_wgrad(c, e); break;
        case EK_WGRAD_T: epi_wgrad_t(c, e); break;
        default: epi_step_end(c, lo); break;
      }
      tc::fence_before();
      fence_async_all();
      epi_bar();
      if (c.tid == 0) st_release(&c.ctl->epi_done, c.sv.base + (uint32_t)k + 1u);
    }
  }
}

struct LaunchP {
  TrainLaunch t;
  const ProgramDev* progs;
  const MemberTc* mtc;
  unsigned char* stash;
  long long stash_bytes;
};

__global__ void __launch_bounds__(kThreadsP, 1) train_tcp_kernel(LaunchP L) {
  extern __shared__ __align__(1024) unsigned char smem[];
  Ctrl* ctl = reinterpret_cast<Ctrl*>(smem + kSmemCtrl);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TrainLaunch& t = L.t;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) { tc::mbar_init(&ctl->full[i], 1); tc::mbar_init(&ctl->empty[i], 1); }
    for (int i = 0; i < 4; ++i) tc::mbar_init(&ctl->accbar[i], 1);
  }
  if (warp == kEpiWarps) tc::tmem_alloc(&ctl->tmem, 512);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = ctl->tmem;
  uint32_t seq = 0, acc_par = 0;
  unsigned char* stash = L.stash + (long long)blockIdx.x * L.stash_bytes;
  const bool dynamic = (int)gridDim.x < t.n_members;
  bool first = true;
  for (;;) {
    if (threadIdx.x == 0) {
      int mi = t.n_members;
      if (dynamic) { mi = atomicAdd(t.work_counter, 1); if (mi < t.n_members) mi = t.order[mi]; else mi = t.n_members; }
      else if (first) mi = blockIdx.x;
      ctl->member = mi;
      ctl->epi_done = 0;
    }
    first = false;
    __syncthreads();
    const int mi = ctl->member;
    if (mi >= t.n_members) break;
    MemberDev& mb = t.members[mi];
    const ProgramDev& pg = L.progs[mb.arch_idx];
    const MemberTc& mt = L.mtc[mi];
    if (warp < kEpiWarps) {
      EpiCtx c;
      c.a = &t.archs[mb.arch_idx]; c.pg = &pg; c.mb = &mb; c.mt = &mt;
      c.smem = smem; c.stash = stash; c.scratch = t.scratch + (long long)blockIdx.x * t.slot_floats; c.ctl = ctl;
      c.tmem = tmem; c.warp = warp; c.lane = lane; c.row = ((warp & 3) << 5) + lane; c.cpart = warp >> 2;
      c.tid = threadIdx.x; c.flags = t.flags;
      c.b1 = mb.beta1; c.b2 = mb.beta2; c.aeps = mb.adam_eps;
      epilogue_role(t, mi, c, acc_par);
    } else if (warp == kEpiWarps) {
      if (lane == 0) mma_role(t, pg, mb, smem, ctl, tmem, seq);
    } else {
      if (lane == 0) producer_role(t, pg, mb, mt, stash, smem, ctl, seq);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int spe = (mb.n_rows + mb.batch - 1) / mb.batch;
      const long long last = mb.steps_done + t.n_steps - 1;
      mb.last_rows = min(mb.batch, mb.n_rows - (int)(last % spe) * mb.batch);
      mb.steps_done += t.n_steps;
      mb.last_slot = blockIdx.x;
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
    }
  }
}

}  // namespace tcp

cudaError_t configure_tcp() {
  return cudaFuncSetAttribute(tcp::train_tcp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::kSmemBytes);
}

cudaError_t launch_xprep(const void* items_dev, int n_items, int max_blocks, cudaStream_t st) {
  if (n_items <= 0) return cudaSuccess;
  dim3 grid(max_blocks < 1 ? 1 : (max_blocks > 64 ? 64 : max_blocks), n_items > 1024 ? 1024 : n_items);
  tcp::xprep_kernel<<<grid, 256, 0, st>>>(static_cast<const tcp::XPrepItem*>(items_dev), n_items);
  return cudaGetLastError();
}

cudaError_t launch_train_tcp(const TrainLaunch& t, const tcp::ProgramDev* progs, const tcp::MemberTc* mtc,
                             unsigned char* stash, long long stash_bytes, int n_sm, cudaStream_t st) {
  const int grid = t.n_members < n_sm ? t.n_members : n_sm;
  if (grid <= 0 || t.n_steps <= 0) return cudaSuccess;
  if (grid < t.n_members) {
    cudaError_t e = cudaMemsetAsync(t.work_counter, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
  }
  tcp::LaunchP L;
  L.t = t; L.progs = progs; L.mtc = mtc; L.stash = stash; L.stash_bytes = stash_bytes;
  tcp::train_tcp_kernel<<<grid, tcp::kThreadsP, tcp::kSmemBytes, st>>>(L);
  return cudaGetLastError();
}

}  // namespace nmb
