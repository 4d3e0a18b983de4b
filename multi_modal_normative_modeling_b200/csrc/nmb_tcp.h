// Pipelined tensor-core training path ("TCP"): host-built step program + device structures.
//
// The fused cVAE step (forward, loss, backward, Adam; cVAE.py:1166-1196 + torch Adam) is a chain of
// small dense contractions.  On B200 the chain is bound by bytes moved between L2 and the SM, so
// this path keeps every matrix in ONE storage format that the tensor core consumes directly --
// BF16 hi/lo planes tiled in the UMMA no-swizzle 8x8 core matrices -- and lets three roles of one
// persistent CTA walk a host-built program:
//   * producer (1 thread): cp.async.bulk (TMA) copies of operand tiles into a 3-slot ring,
//   * MMA      (1 thread): tcgen05.mma, 3 passes (hi*hi, lo*hi, hi*lo) per K=16 step into TMEM,
//   * epilogue (8 warps) : tcgen05.ld -> activation / loss / leaky-relu' / Adam -> planes written to
//                          shared memory (next operand, no staging pass) and to the backward stash.
// The minibatch is processed as two independent 128-row halves so that the epilogue of one half
// overlaps the MMAs of the other.
//
// Canonical block (R rows, cg column groups of 8):   byte(r, c, plane) =
//        (c >> 3) * 32R  +  plane * 16R  +  r * 16  +  (c & 7) * 2
// The same bytes serve as a K-major operand (MN index = r, K index = c: SBO = 128, LBO = 32R) and as
// an MN-major operand (MN index = c, K index = r: SBO = 32R, LBO = 128), so activations, output
// gradients and weights are each stored once and used by forward, dgrad and wgrad.
#pragma once
#include <cstdlib>
#include <map>
#include <vector>

#include "nmb_common.cuh"

namespace nmb {
namespace tcp {

constexpr int kGroups = 3;                        // epilogue groups: minibatch half 0, half 1, optimiser (Adam)
constexpr int kGroupWarps = 4;                    // one warp per TMEM lane quadrant
constexpr int kGroupThreads = kGroupWarps * 32;
constexpr int kEpiWarps = kGroups * kGroupWarps;  // 12 + MMA + producer = 14 warps -> 128 registers per thread
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreadsP = kEpiThreads + 64;        // + MMA warp + producer warp
constexpr int kActBytes = 65536;                   // one 128-row operand block, up to 128 columns (hi + lo)
constexpr int kSlotBytes = 32768;                  // one ring tile
constexpr int kSlots = 3;
constexpr int kSmemRing = 2 * kActBytes;
constexpr int kSmemCtrl = kSmemRing + kSlots * kSlotBytes;
constexpr int kSmemBytes = kSmemCtrl + 1024;       // 230 400 B <= 227 KB
constexpr int kAcc0 = 0, kWacc0 = 256;             // TMEM columns: acc[h] = 128 h, wacc[b] = 256 + 128 b

enum Space : int { SP_NONE = 0, SP_W = 1, SP_STASH = 2, SP_X = 3 };

enum EpiKind : int {
  EK_HIDDEN = 0,      // leaky-relu, planes -> ACT[h] + stash
  EK_HEAD,            // [mu | logvar] fp32 -> mulv array
  EK_LATENT,          // fusion + reparameterisation + KL; decoder inputs [z | c | 1]
  EK_COPY,            // stash block -> ACT[h]
  EK_RECON,           // loss terms + d/dx_recon planes -> stash; logvar_out gradient partials
  EK_LAM,             // Adam on logvar_out
  EK_DGRAD,           // leaky-relu' -> planes -> ACT[h]
  EK_DZ,              // accumulate d/dz
  EK_LATENT_BWD,      // latent + fusion backward -> d[mu | logvar] planes
  EK_WGRAD,           // weight gradient (rows = out) fused with Adam; planes of the new weights
  EK_WGRAD_T,         // transposed weight gradient of decoder_mean_layer (rows = in)
  EK_STEP_END,        // loss reduction, alpha (gPoE) update; publishes the new weight planes to the TMA proxy
  EK_FENCE,           // publishes every stash block written so far to the TMA proxy (start of the backward pass)
  EK_HEAD_LATENT,     // one modality, Z <= 16: head accumulator -> reparameterisation -> [z | c | 1] planes, in registers
  EK_DZ_LATENT_BWD,   // one modality, Z <= 16: d/dz accumulator -> d[mu | logvar] planes, in registers
  EK_XHAT             // forward-only program (pred_recon): x_recon tile -> fp32 rows of the caller's output
};

// One ring tile (+ optional A tile) and the MMAs issued on it.
struct alignas(16) Step {
  long long b_off, a_off;           // byte offsets inside their spaces
  unsigned b_bytes, a_bytes;        // a_bytes == 0: A is resident in ACT[half]
  int dep;                          // joint epilogue item (index + 1, this step) that published the tile data; 0 = none
  int mma_dep;                      // item of this half's epilogue group that must finish before these MMAs issue
  int mma_dep_joint;                // item of the optimiser group (Adam epilogue) that must finish before these MMAs issue
  unsigned a_start;                 // byte offset inside ACT[half]
  unsigned a_lbo, a_sbo, a_kadv, a_lo;
  unsigned b_lbo, b_sbo, b_kadv, b_lo;
  unsigned short ksteps, n, tmem_col;
  unsigned char half, b_space, a_space, x_mod;
  unsigned char a_mn, b_mn;         // 1 = MN-major
  unsigned char first;              // first MMA overwrites the accumulator
  unsigned char commit;             // 0 none, 1 always, 2 only when half 1 is inactive
  unsigned char commit_buf;         // accumulator barrier 0..3
  unsigned char commit2, commit2_buf;   // optional second commit of the same step (same encoding)
  unsigned char dep_grp;            // epilogue group whose counter `dep` refers to (0 / 1; 2 = both)
  unsigned char a_hold;             // the A tile's ring slot is kept for the next step ...
  unsigned char b_held;             // ... which reads it as its B operand (no tile of its own) and releases it
  unsigned char b2;                 // two-part step: the FIRST ring tile (a_space / a_off / a_bytes) is B of part 1 (k_first
                                    // K steps), the second tile B of part 2 (the rest); A stays resident in ACT[half]
  unsigned char k_first;            // K steps of part 1
  unsigned char p2_new;             // part 2 is a new N chunk (A restarts at a_start, accumulator columns tmem2, width n2,
                                    // overwrite like part 1) instead of the continuation of part 1 along K
  unsigned short n2, tmem2;
};

// Compact copy of the MMA-relevant fields of a Step.  The table travels in the KERNEL PARAMETERS (constant
// bank), so the MMA warp reads it with uniform loads and issues tcgen05.mma straight from uniform registers.
struct MStep {
  unsigned a_start;                                   // byte offset inside ACT[half] (A resident)
  unsigned short a_lbo, a_sbo, a_kadv, a_lo;          // descriptor strides in 16-byte units
  unsigned short b_lbo, b_sbo, b_kadv, b_lo;
  unsigned short ksteps, n, tmem_col;
  short mma_dep, mma_dep_joint;
  unsigned char half, a_tile, a_mn, b_mn, first, commit, commit_buf, commit2;   // commit2_buf = accumulator of `half`
  unsigned char k_first, p2_new;                      // two-part steps (a_tile bit 2), see Step
  unsigned short n2, tmem2;
};
constexpr int kMaxParamSteps = 400;                   // 44 B each: 17.6 KB of the 32 KB parameter space
constexpr int kMaxParamArchs = 8;

inline MStep to_mstep(const Step& s) {
  MStep m{};
  m.a_start = s.a_start;
  m.a_lbo = (unsigned short)(s.a_lbo >> 4); m.a_sbo = (unsigned short)(s.a_sbo >> 4);
  m.a_kadv = (unsigned short)(s.a_kadv >> 4); m.a_lo = (unsigned short)(s.a_lo >> 4);
  m.b_lbo = (unsigned short)(s.b_lbo >> 4); m.b_sbo = (unsigned short)(s.b_sbo >> 4);
  m.b_kadv = (unsigned short)(s.b_kadv >> 4); m.b_lo = (unsigned short)(s.b_lo >> 4);
  m.ksteps = s.ksteps; m.n = s.n; m.tmem_col = s.tmem_col;
  m.mma_dep = (short)s.mma_dep; m.mma_dep_joint = (short)s.mma_dep_joint;
  m.half = s.half; m.a_tile = (unsigned char)((s.b2 ? 4 : (s.a_bytes ? 1 : 0)) | (s.a_hold ? 2 : 0)); m.a_mn = s.a_mn;
  m.b_mn = (unsigned char)(s.b_mn | (s.b_held ? 2 : 0));      // a_tile bit 1: keep the A slot; b_mn bit 1: B = kept slot
  m.first = s.first; m.commit = s.commit; m.commit_buf = s.commit_buf; m.commit2 = s.commit2;
  m.k_first = s.b2 ? (s.k_first ? s.k_first : (unsigned char)(s.ksteps / 2)) : 0; m.p2_new = s.p2_new; m.n2 = s.n2; m.tmem2 = s.tmem2;
  return m;
}

struct alignas(16) Epi {
  int kind, half, buf, mod;         // half: 0/1, 2 = joint.  buf: accumulator barrier to wait on, -1 = none
  int n_mma, n_valid, n_cols;       // MMA N; valid output columns; columns written (multiple of 16)
  int tmem_col;                     // absolute TMEM column of the accumulator
  int col0;                         // first logical column / row handled by this item (chunked ops)
  int to_act;                       // planes also written to ACT[half]
  long long stash_off;              // destination block (bytes, stash space), -1 = none
  long long src_off;                // EK_COPY source / EK_DGRAD sign source (stash space), -1 = none
  int src_cg;                       // EK_COPY: column groups to copy
  // parameters (EK_WGRAD / EK_WGRAD_T / EK_LAM)
  long long p_off; int p_ld, p_rows, p_cols;   // float offset of the augmented matrix, row stride, out, in + 1
  long long wp_off; int wp_R;       // weight planes block (byte offset in SP_W) and its R
  long long mst_off; int mst_R;     // Adam master state of the layer (float offset in the slot's master buffer, lane extent)
  int row0;                         // EK_WGRAD: first weight row of this planes block (out-layer blocks)
  int last;                         // EK_RECON: last tile of the modality (lam partials complete)
  int split_all;                    // Adam item after the last per-half item of the step: all three groups share it
  int wait_optim;                   // split_all: last optimiser-only item (index + 1) an activation group must see finished
                                    // before it trusts its view of the accumulator barrier phases
};

// Compact copy of an Epi (64 B).  Like the MMA step table it travels in the KERNEL PARAMETERS, so that decoding an
// item costs a few uniform loads from the constant bank instead of a 144-byte read through L2 on the critical path
// of every item.  Architectures whose items do not fit (kMaxParamEpis in total) read the global-memory table.
struct alignas(16) EpiP {
  int stash_off, src_off, p_off, wp_off, mst_off;       // byte / float offsets (all < 2^31), -1 = none
  unsigned short n_mma, n_valid, n_cols, tmem_col, col0, src_cg, p_ld, p_rows, p_cols, wp_R, mst_R, row0, wait_optim;
  signed char kind, half, buf, mod;
  unsigned char to_act, last, split_all, pad_;
  // Per epilogue group g: the next item the group executes after this one (255 = none) and, for THIS item, the
  // weight-gradient accumulator barriers (bit 0 = wacc[0], bit 1 = wacc[1]) whose phase the group must flip before it,
  // because it skipped optimiser-only items that consumed them.  Lets a group jump from own item to own item instead of
  // decoding every item of the other groups on its critical path.
  unsigned char next_own[3], flip_before[3];
  unsigned char pad3_[2];
};
static_assert(sizeof(EpiP) == 64, "EpiP must stay 64 bytes");
constexpr int kMaxParamEpis = 160;                    // 10 KB

// which group(s) execute an item (all rows of the minibatch present)
inline bool epi_owned_by(const Epi& e, int g) {
  if (e.kind == EK_STEP_END || e.split_all) return true;
  if (e.kind == EK_WGRAD || e.kind == EK_WGRAD_T) return g == 2;
  return e.half == g;
}
// fill next_own / flip_before of a program's compact items; first_own[g] = the group's first item
inline void link_group_items(const std::vector<Epi>& epis, EpiP* out, unsigned char first_own[3], unsigned char first_flip[3]) {
  const int n = (int)epis.size();
  for (int g = 0; g < 3; ++g) {
    int prev = -1;
    unsigned flip = 0;
    first_own[g] = 255; first_flip[g] = 0;
    for (int k = 0; k < n; ++k) {
      const Epi& e = epis[k];
      if (!epi_owned_by(e, g)) {
        if ((e.kind == EK_WGRAD || e.kind == EK_WGRAD_T) && e.buf >= 2) flip ^= 1u << (e.buf - 2);
        continue;
      }
      if (prev < 0) { first_own[g] = (unsigned char)k; first_flip[g] = (unsigned char)flip; }
      else out[prev].next_own[g] = (unsigned char)k;
      out[k].flip_before[g] = (unsigned char)flip;
      out[k].next_own[g] = 255;
      flip = 0;
      prev = k;
    }
  }
}

inline bool epip_fits(const Epi& e) {
  return e.stash_off < (1LL << 31) && e.src_off < (1LL << 31) && e.p_off < (1LL << 31) && e.wp_off < (1LL << 31) &&
         e.mst_off < (1LL << 31) && e.p_ld < 65536 && e.p_rows < 65536 && e.p_cols < 65536;
}
inline EpiP to_epip(const Epi& e) {
  EpiP q{};
  q.stash_off = (int)e.stash_off; q.src_off = (int)e.src_off; q.p_off = (int)e.p_off; q.wp_off = (int)e.wp_off;
  q.mst_off = (int)e.mst_off;
  q.n_mma = (unsigned short)e.n_mma; q.n_valid = (unsigned short)e.n_valid; q.n_cols = (unsigned short)e.n_cols;
  q.tmem_col = (unsigned short)e.tmem_col; q.col0 = (unsigned short)e.col0; q.src_cg = (unsigned short)e.src_cg;
  q.p_ld = (unsigned short)e.p_ld; q.p_rows = (unsigned short)e.p_rows; q.p_cols = (unsigned short)e.p_cols;
  q.wp_R = (unsigned short)e.wp_R; q.mst_R = (unsigned short)e.mst_R; q.row0 = (unsigned short)e.row0;
  q.wait_optim = (unsigned short)e.wait_optim;
  q.kind = (signed char)e.kind; q.half = (signed char)e.half; q.buf = (signed char)e.buf; q.mod = (signed char)e.mod;
  q.to_act = (unsigned char)e.to_act; q.last = (unsigned char)e.last; q.split_all = (unsigned char)e.split_all;
  return q;
}
__host__ __device__ inline Epi from_epip(const EpiP& q) {
  Epi e;
  e.kind = q.kind; e.half = q.half; e.buf = q.buf; e.mod = q.mod;
  e.n_mma = q.n_mma; e.n_valid = q.n_valid; e.n_cols = q.n_cols; e.tmem_col = q.tmem_col; e.col0 = q.col0;
  e.to_act = q.to_act; e.stash_off = q.stash_off; e.src_off = q.src_off; e.src_cg = q.src_cg;
  e.p_off = q.p_off; e.p_ld = q.p_ld; e.p_rows = q.p_rows; e.p_cols = q.p_cols;
  e.wp_off = q.wp_off; e.wp_R = q.wp_R; e.mst_off = q.mst_off; e.mst_R = q.mst_R;
  e.row0 = q.row0; e.last = q.last; e.split_all = q.split_all; e.wait_optim = q.wait_optim;
  return e;
}

// Weight block of the per-member planes buffer (prologue conversion fp32 -> planes).
struct WBlock {
  long long p_off; int p_ld, row0, rows_valid, cols_valid;
  long long wp_off; int R, cg;
  int transposed;     // 1: planes block rows = INPUT index (R = round16(in + 1)), column groups = output rows (cg groups of 8)
};

// Adam master state of one linear layer inside the per-slot master buffer.  While a member is resident its
// fp32 parameters and moments live in a lane-major layout so that the Adam epilogues are coalesced:
//   kind 0 (lane = output row o):  float index = ((i >> 2) * R + o) * 4 + (i & 3)
//   kind 1 (lane = input index i): float index = ((o >> 2) * R + i) * 4 + (o & 3)     (decoder_mean_layer)
// They are gathered from / scattered back to the caller's row-major buffers at member start / end.
struct MLayer {
  long long p_off; int p_ld, rows, cols;   // augmented matrix [rows][p_ld], cols = in + 1
  long long mst_off; int R, kind;
};

// fp32 side arrays inside the stash slot (byte offsets) and their strides (floats).
struct Layout {
  long long mulv[NMB_MAX_MOD]; int ld_mulv;      // [256][ld_mulv] heads per modality
  long long g0[NMB_MAX_MOD][2], dmulv[NMB_MAX_MOD][2];   // decoder-input / head-gradient blocks per half
  int wout_cg[NMB_MAX_MOD];                       // column groups of a decoder_mean_layer planes block
  long long zbuf, dz;                             // [256][Z]
  long long lampart[NMB_MAX_MOD];                 // [8][round4(D)] logvar_out gradient partials
  long long ivtab[NMB_MAX_MOD];                   // [round4(D) + 8] exp(-logvar_out), zero padded
  long long dxh_blk[NMB_MAX_MOD];                 // first d/dx_recon block of half 0 (blocks of 32 KB, [half][j])
  int n_dxh_blk[NMB_MAX_MOD];
  long long stash_bytes;
  long long wplanes_bytes;
  long long master_floats;                        // per moment; the slot holds 3 of them (p, m, v)
  int x_cg[NMB_MAX_MOD];                          // column groups of a dataset block
  int x_quads[NMB_MAX_MOD];                       // 4-column groups of a lane-major target block
  int c_cg;                                       // column groups of a decoder-input template block
};

struct Program {
  std::vector<Step> steps;
  std::vector<Epi> epis;
  std::vector<WBlock> wblocks;
  std::vector<MLayer> mlayers;
  Layout lay;
  bool eligible = false;
};

struct ProgramDev {          // per architecture, device pointers
  const Step* steps; const Epi* epis; const WBlock* wblocks; const MLayer* mlayers;
  int n_steps, n_epis, n_wblocks, n_mlayers;
  Layout lay;
};

struct MemberTc {            // per member
  unsigned char* wplanes;
  const unsigned char* xplanes[NMB_MAX_MOD];   // dataset blocks [pos][half], 128 rows each
  const unsigned char* cplanes[NMB_MAX_MOD];   // decoder-input templates [0 (Z) | c | 1] per block, canonical planes
  const float* xlm[NMB_MAX_MOD];               // ROI targets per block, lane-major fp32: [(col>>2)][row][4]
  int n_half;                                  // halves per minibatch = ceil(batch / 128)
};

// Per-member arguments of a forward-only (reconstruction) launch: rows to reconstruct and where the results go.
struct ReconTc {
  const unsigned char* xplanes[NMB_MAX_MOD];   // [tile][half] canonical blocks of the rows (xprep_kernel, batch 256)
  const unsigned char* cplanes[NMB_MAX_MOD];   // decoder-input templates of the rows
  const float* xc[NMB_MAX_MOD];                // the packed fp32 rows themselves (covariates of the unfused latent path)
  float* xhat[NMB_MAX_MOD];                    // out [n_rows][D_m] or NULL
  float* mu; float* logvar;                    // out [n_rows][Z] or NULL
  const float* eps;                            // injected draws [n_rows][Z] or NULL (Philox stream 1)
  int n_rows;
};
struct ReconWork { int member, tile0, n_tiles, rt; };   // tiles of 256 rows; rt = index of the row set's ReconTc

// One dataset (packed fp32 rows of one modality) to be re-tiled into 128-row blocks per (minibatch, half).
struct XPrepItem {
  const float* xc; unsigned char* out; unsigned char* cplanes; float* xlm;
  int n_rows, batch, ldx, k_valid, cg, n_half, d, c_dim, z, c_cg, quads;
};

__host__ __device__ inline int round16(int v) { return (v + 15) & ~15; }

// ---- host: build the step program of one architecture ------------------------------------------
// fwd_only: the forward half only, ending in EK_XHAT items (same weight-plane and stash layout as the training program,
// so both programs run on the same per-member planes); nothing is stashed for a backward pass.
inline Program build_program(const ArchDesc& a, bool fwd_only = false) {
  Program P;
  const int M = a.M, L = a.L, Z = a.Z, C = a.C;
  if (2 * Z > 128 || Z + C + 1 > 128) return P;
  if (a.head_kind || a.family) return P;            // supervised heads and the baseline families run on the generic engines
  for (int l = 0; l < L; ++l) if (a.hidden[l] > 127) return P;
  P.eligible = true;
  Layout& lay = P.lay;
  lay = Layout{};
  // dev knob NMB_TCP_MERGE (bit mask, default 7): two-part MMA steps for 1 = the K-split weight gradients, 2 = the
  // two-tile forward layers, 4 = the two-chunk data gradients; 0 = one MMA step per ring tile everywhere
  const char* mk_env = getenv("NMB_TCP_MERGE");
  const int merge_mask = mk_env ? atoi(mk_env) : 7;
  const bool merge_k = (merge_mask & 1) != 0, merge_fwd = (merge_mask & 2) != 0, merge_dg = (merge_mask & 4) != 0;
  const bool split_recon = (merge_mask & 8) != 0;   // 8: x_recon tiles alternate between two accumulator halves (measured: no gain, off)

  // ---- stash + planes layout ----
  long long so = 0, wo = 0;
  auto alloc = [&](long long bytes) { long long r = so; so += (bytes + 127) & ~127LL; return r; };
  auto act_block = [&](int cols) { return alloc((long long)(round16(cols) / 8) * 4096); };
  struct WRef { long long wp_off; int R, cg, row0, rows; };
  auto wblock = [&](const LinDesc& w, int row0, int rows, int R) {
    WBlock b; b.p_off = w.off; b.p_ld = w.ld; b.row0 = row0; b.rows_valid = rows; b.cols_valid = w.in + 1;
    b.R = R; b.cg = round16(w.in + 1) / 8; b.wp_off = wo; b.transposed = 0;
    wo += (long long)b.cg * 32 * R;
    P.wblocks.push_back(b);
    return WRef{b.wp_off, R, b.cg, row0, rows};
  };
  long long mo = 0;
  std::map<long long, MLayer> mst;   // keyed by the layer's parameter offset
  auto master = [&](const LinDesc& w, int kind) {
    MLayer ml; ml.p_off = w.off; ml.p_ld = w.ld; ml.rows = w.out; ml.cols = w.in + 1; ml.kind = kind;
    const int lanes = kind == 0 ? w.out : w.in + 1, other = kind == 0 ? w.in + 1 : w.out;
    ml.R = (lanes + 31) & ~31;
    ml.mst_off = mo;
    mo += (long long)((other + 3) / 4) * ml.R * 4;
    P.mlayers.push_back(ml);
    mst[w.off] = ml;
  };
  // per-half activation blocks: [m][h]
  std::vector<std::vector<long long>> s_h(M), s_k(M);     // [m][l*2+h]
  std::vector<long long> s_g0(M * 2), s_dmulv(M * 2);
  std::vector<std::vector<WRef>> w_enc(M), w_dec(M), w_out(M);
  std::vector<WRef> w_head(M);
  lay.ld_mulv = round4(2 * Z);
  lay.c_cg = round16(Z + C + 1) / 8;
  for (int m = 0; m < M; ++m) {
    const ModDesc& q = a.mod[m];
    s_h[m].resize(L * 2); s_k[m].resize(L * 2);
    for (int l = 0; l < L; ++l)
      for (int h = 0; h < 2; ++h) {
        s_h[m][l * 2 + h] = act_block(a.hidden[l] + 1);
        s_k[m][l * 2 + h] = act_block(a.hidden[L - 1 - l] + 1);
      }
    for (int h = 0; h < 2; ++h) {
      s_g0[m * 2 + h] = act_block(Z + C + 1); s_dmulv[m * 2 + h] = act_block(2 * Z);
      lay.g0[m][h] = s_g0[m * 2 + h]; lay.dmulv[m][h] = s_dmulv[m * 2 + h];
    }
    lay.wout_cg[m] = round16(q.outl.in + 1) / 8;
    // decoder_mean_layer: with one modality and D <= 128 its output gradient stays in ACT[h] (regular layer);
    // otherwise d/dx_recon goes through 64-column stash blocks and the weight gradient is transposed
    const bool fast_out = M == 1 && q.D <= 128;
    lay.n_dxh_blk[m] = fast_out ? 0 : (q.D + 63) / 64;
    lay.dxh_blk[m] = alloc((long long)2 * lay.n_dxh_blk[m] * 32768 + 128);
    lay.mulv[m] = alloc((long long)256 * lay.ld_mulv * 4);
    lay.lampart[m] = alloc((long long)8 * round4(q.D) * 4);
    lay.ivtab[m] = alloc((long long)(round4(q.D) + 8) * 4);
    lay.x_cg[m] = round16(q.D + C + 1) / 8;
    lay.x_quads[m] = (q.D + 3) / 4;
    for (int l = 0; l < L; ++l) { master(q.enc[l], 0); master(q.dec[l], 0); }
    master(q.head, 0); master(q.outl, fast_out ? 0 : 1);
    for (int l = 0; l < L; ++l) w_enc[m].push_back(wblock(q.enc[l], 0, q.enc[l].out, round16(q.enc[l].out)));
    w_head[m] = wblock(q.head, 0, q.head.out, round16(q.head.out));
    for (int l = 0; l < L; ++l) w_dec[m].push_back(wblock(q.dec[l], 0, q.dec[l].out, round16(q.dec[l].out)));
    if (fast_out) w_out[m].push_back(wblock(q.outl, 0, q.D, round16(q.D)));
    else {
      // wide / multi-modality output layer: ONE block stored transposed (rows = input index, 8-column groups = output
      // rows), so that the transposed weight-gradient epilogue (lane = input index) rewrites it with 16-byte stores.
      // Any run of 8 groups (64 output rows) is a contiguous tile: MN-major B of the forward GEMM, K-major B of dgrad.
      WBlock b; b.p_off = q.outl.off; b.p_ld = q.outl.ld; b.row0 = 0; b.rows_valid = q.D; b.cols_valid = q.outl.in + 1;
      b.R = round16(q.outl.in + 1); b.cg = 8 * lay.n_dxh_blk[m]; b.wp_off = wo; b.transposed = 1;
      wo += (long long)b.cg * 32 * b.R;
      P.wblocks.push_back(b);
      w_out[m].push_back(WRef{b.wp_off, b.R, b.cg, 0, q.D});
    }
  }
  lay.zbuf = alloc((long long)256 * Z * 4);
  lay.dz = alloc((long long)256 * Z * 4);
  lay.stash_bytes = so;
  lay.wplanes_bytes = (wo + 127) & ~127LL;
  lay.master_floats = (mo + 31) & ~31LL;

  // ---- program ----
  int act_ready[2] = {0, 0};    // epilogue item (index+1) after which ACT[h] holds what the next MMA group reads
  int acc_free[4] = {0, 0, 0, 0};
  auto accbuf = [](int h) { return h; };
  std::map<long long, int> ready;   // stash block -> epilogue item (index + 1) that wrote it
  auto push_epi = [&](Epi e) {      // returns index + 1
    P.epis.push_back(e);
    const int id = (int)P.epis.size();
    if (e.stash_off >= 0) ready[e.stash_off] = id;
    return id;
  };
  auto new_epi = [&](int kind, int half, int buf, int mod) {
    Epi e{}; e.kind = kind; e.half = half; e.buf = buf; e.mod = mod; e.stash_off = -1; e.src_off = -1;
    // accumulator barriers: 0 / 1 = acc[h], 2 / 3 = wacc[b], 4 / 5 = the upper 64 columns of acc[h] used as a second
    // x_recon tile buffer (the tiles of the wide output layer alternate between the two halves of acc[h])
    e.tmem_col = buf < 0 ? 0 : (buf < 2 ? kAcc0 + 128 * buf : (buf < 4 ? kWacc0 + 128 * (buf - 2) : kAcc0 + 128 * (buf - 4) + 64));
    return e;
  };
  // MMA-issue dependencies: items of the step's own half go to that group's counter, joint items to both
  auto need = [&](Step& s, int id) {
    if (id <= 0) return;
    const int kind = P.epis[id - 1].kind;
    if (kind == EK_WGRAD || kind == EK_WGRAD_T) { if (id > s.mma_dep_joint) s.mma_dep_joint = id; }
    else if (id > s.mma_dep) s.mma_dep = id;
  };
  auto base_step = [&](int h) {
    Step s{}; s.half = (unsigned char)h; s.b_space = SP_NONE; s.a_space = SP_NONE;
    return s;
  };
  auto set_a_kmajor = [](Step& s, unsigned start) {   // 128-row block (ACT or ring), K-major
    s.a_mn = 0; s.a_start = start; s.a_lbo = 4096; s.a_sbo = 128; s.a_kadv = 8192; s.a_lo = 2048;
  };
  auto set_a_mnmajor = [](Step& s) {                  // ACT[h] as MN-major A (M = columns), K = 128 rows
    s.a_mn = 1; s.a_start = 0; s.a_sbo = 4096; s.a_lbo = 128; s.a_kadv = 256; s.a_lo = 2048;
  };
  auto set_b = [](Step& s, int space, long long off, int R, int g0, int ng, bool mn) {
    s.b_space = (unsigned char)space; s.b_off = off + (long long)g0 * 32 * R; s.b_bytes = (unsigned)(ng * 32 * R);
    s.b_mn = mn ? 1 : 0; s.b_lo = 16 * R;
    if (mn) { s.b_sbo = 32 * R; s.b_lbo = 128; s.b_kadv = 256; }
    else { s.b_lbo = 32 * R; s.b_sbo = 128; s.b_kadv = 64 * R; }
  };
  auto tile_groups = [](int R) { int g = 1024 / R; return g > 16 ? 16 : (g < 2 ? 2 : g & ~1); };
  // One modality: a stashed hidden activation of 9 .. 16 column groups is stored as two 64-row sub-blocks (the
  // weight gradient that reads it is then split over K, see emit_wgrad_part); `cols` = layer width + 1.
  auto is64 = [&](int cols) { const int cg = round16(cols) / 8; return M == 1 && cg > 8 && cg <= 16; };

  // forward-type group: acc[h][0..N) = sum_k A[.,k] B[n,k]; A from ACT[h] or ring tiles (a_space/a_base)
  auto emit_fwd = [&](int h, const WRef& w, int a_space, long long a_base, int a_dep, int x_mod) {
    const int buf = accbuf(h);
    const int tg = tile_groups(w.R) > 8 && a_space != SP_NONE ? 8 : tile_groups(w.R);
    if (merge_fwd && a_space == SP_NONE && w.cg > tg && w.cg <= 2 * tg) {
      // both K tiles of the weight block in ONE MMA step (A resident in ACT[h], continuing along K)
      Step s = base_step(h);
      set_b(s, SP_W, w.wp_off, w.R, tg, w.cg - tg, false);                       // part 2: groups tg .. cg
      s.b2 = 1; s.k_first = (unsigned char)(tg / 2);
      s.a_space = SP_W; s.a_off = w.wp_off; s.a_bytes = (unsigned)(tg * 32 * w.R);   // part 1: groups 0 .. tg
      set_a_kmajor(s, 0);
      s.dep = a_dep;
      s.ksteps = (unsigned short)(w.cg / 2); s.n = (unsigned short)w.R; s.tmem_col = (unsigned short)(kAcc0 + 128 * buf);
      s.first = 1;
      need(s, acc_free[buf]); need(s, act_ready[h]);
      s.commit = 1; s.commit_buf = (unsigned char)buf;
      P.steps.push_back(s);
      return;
    }
    for (int g0 = 0, j = 0; g0 < w.cg; g0 += tg, ++j) {
      const int ng = w.cg - g0 < tg ? w.cg - g0 : tg;
      Step s = base_step(h);
      set_b(s, SP_W, w.wp_off, w.R, g0, ng, false);
      if (a_space == SP_NONE) set_a_kmajor(s, (unsigned)g0 * 4096);
      else {
        set_a_kmajor(s, 0);
        s.a_space = (unsigned char)a_space; s.a_off = a_base + (long long)g0 * 4096; s.a_bytes = (unsigned)ng * 4096;
        s.x_mod = (unsigned char)x_mod;
      }
      s.dep = a_dep;
      s.ksteps = (unsigned short)(ng / 2); s.n = (unsigned short)w.R; s.tmem_col = (unsigned short)(kAcc0 + 128 * buf);
      s.first = j == 0;
      if (j == 0) { need(s, acc_free[buf]); if (a_space == SP_NONE) need(s, act_ready[h]); }
      s.commit = g0 + tg >= w.cg ? 1 : 0; s.commit_buf = (unsigned char)buf;
      P.steps.push_back(s);
    }
  };
  // dgrad-type group: acc[h][i] = sum_o ACT[h][., o] W[o][i], N-chunks of the weight block
  auto emit_dgrad = [&](int h, const WRef& w, int n_need, bool commit) {
    const int buf = accbuf(h);
    const int tg = tile_groups(w.R);
    const int cg = (round16(n_need) / 8) < w.cg ? round16(n_need) / 8 : w.cg;
    if (merge_dg && cg > tg && cg <= 2 * tg && 2 * (w.R / 16) < 256) {
      // both N chunks in ONE MMA step: part 1 -> columns 0 .. 8 tg, part 2 (new N chunk, A restarts) -> the rest
      Step s = base_step(h);
      set_b(s, SP_W, w.wp_off, w.R, tg, cg - tg, true);                           // part 2
      s.b2 = 1; s.k_first = (unsigned char)(w.R / 16); s.p2_new = 1;
      s.a_space = SP_W; s.a_off = w.wp_off; s.a_bytes = (unsigned)(tg * 32 * w.R);   // part 1: groups 0 .. tg
      set_a_kmajor(s, 0);
      s.ksteps = (unsigned short)(2 * (w.R / 16)); s.n = (unsigned short)(tg * 8);
      s.tmem_col = (unsigned short)(kAcc0 + 128 * buf);
      s.n2 = (unsigned short)((cg - tg) * 8); s.tmem2 = (unsigned short)(kAcc0 + 128 * buf + tg * 8);
      s.first = 1;
      need(s, act_ready[h]); need(s, acc_free[buf]);
      s.commit = commit ? 1 : 0; s.commit_buf = (unsigned char)buf;
      P.steps.push_back(s);
      return;
    }
    for (int g0 = 0, j = 0; g0 < cg; g0 += tg, ++j) {
      const int ng = cg - g0 < tg ? cg - g0 : tg;
      Step s = base_step(h);
      set_b(s, SP_W, w.wp_off, w.R, g0, ng, true);
      set_a_kmajor(s, 0);
      s.ksteps = (unsigned short)(w.R / 16); s.n = (unsigned short)(ng * 8);
      s.tmem_col = (unsigned short)(kAcc0 + 128 * buf + g0 * 8);
      s.first = 1;
      if (j == 0) { need(s, act_ready[h]); need(s, acc_free[buf]); }
      s.commit = (commit && g0 + tg >= cg) ? 1 : 0; s.commit_buf = (unsigned char)buf;
      P.steps.push_back(s);
    }
  };
  // wgrad part for half h: wacc[wb][.., 64 j + ..) += ACT[h]^T (MN-major A) x B tiles (MN-major, N-chunks)
  // acc_commit: the part also commits acc[h] on its last step -- the preceding dgrad group of the same half
  // left it open so that the epilogue overwriting ACT[h] cannot start before these MMAs have read ACT[h].
  auto emit_wgrad_part = [&](int h, int wb, int b_space, long long b_base, int g_first, int g_count, int x_mod,
                             bool acc_commit, int blk64_cg) {
    const int b_dep = b_space == SP_STASH ? ready[b_base] : 0;
    if (blk64_cg && merge_k) {
      // both 64-row sub-blocks of B in ONE MMA step (two ring tiles, 8 K steps): the backward pass is bound by the
      // per-step cost of the MMA issuer, and this saves one step per (layer, half)
      Step s = base_step(h);
      s.b2 = 1;
      s.a_space = (unsigned char)b_space; s.a_off = b_base; s.a_bytes = (unsigned)blk64_cg * 2048;
      s.b_space = (unsigned char)b_space; s.b_off = b_base + (long long)blk64_cg * 2048; s.b_bytes = (unsigned)blk64_cg * 2048;
      s.b_mn = 1; s.b_lo = 1024; s.b_sbo = 2048; s.b_lbo = 128; s.b_kadv = 256;
      set_a_mnmajor(s); s.a_start = 0;
      s.dep = b_dep;
      s.ksteps = 8; s.n = (unsigned short)(g_count * 8); s.tmem_col = (unsigned short)(kWacc0 + 128 * wb);
      s.first = h == 0;
      need(s, act_ready[h]); need(s, acc_free[2 + wb]);
      s.commit = h == 1 ? 1 : 2;
      s.commit_buf = (unsigned char)(2 + wb);
      if (acc_commit) { s.commit2 = 1; s.commit2_buf = (unsigned char)accbuf(h); }
      P.steps.push_back(s);
      return;
    }
    if (blk64_cg) {
      // B block stored as two 64-row sub-blocks (all column groups contiguous, <= 32 KB each): the GEMM is split over K
      // (batch rows) instead of N, so that the MN-major A operand (ACT[h]^T) is read once per pass, not once per N chunk
      for (int kh = 0; kh < 2; ++kh) {
        Step s = base_step(h);
        s.b_space = (unsigned char)b_space; s.b_off = b_base + (long long)kh * blk64_cg * 2048; s.b_bytes = (unsigned)blk64_cg * 2048;
        s.b_mn = 1; s.b_lo = 1024; s.b_sbo = 2048; s.b_lbo = 128; s.b_kadv = 256;
        set_a_mnmajor(s); s.a_start = (unsigned)kh * 1024;
        s.dep = b_dep;
        s.ksteps = 4; s.n = (unsigned short)(g_count * 8); s.tmem_col = (unsigned short)(kWacc0 + 128 * wb);
        s.first = h == 0 && kh == 0;
        if (kh == 0) { need(s, act_ready[h]); need(s, acc_free[2 + wb]); }
        const bool last = kh == 1;
        s.commit = !last ? 0 : (h == 1 ? 1 : 2);
        s.commit_buf = (unsigned char)(2 + wb);
        if (last && acc_commit) { s.commit2 = 1; s.commit2_buf = (unsigned char)accbuf(h); }
        P.steps.push_back(s);
      }
      return;
    }
    for (int g0 = 0; g0 < g_count; g0 += 8) {
      const int ng = g_count - g0 < 8 ? g_count - g0 : 8;
      Step s = base_step(h);
      set_b(s, b_space, b_base, 128, g_first + g0, ng, true);
      s.x_mod = (unsigned char)x_mod;
      set_a_mnmajor(s);
      s.dep = b_dep;
      s.ksteps = 8; s.n = (unsigned short)(ng * 8); s.tmem_col = (unsigned short)(kWacc0 + 128 * wb + g0 * 8);
      s.first = h == 0;
      if (g0 == 0) { need(s, act_ready[h]); need(s, acc_free[2 + wb]); }
      const bool last = g0 + 8 >= g_count;
      s.commit = !last ? 0 : (h == 1 ? 1 : 2);
      s.commit_buf = (unsigned char)(2 + wb);
      if (last && acc_commit) { s.commit2 = 1; s.commit2_buf = (unsigned char)accbuf(h); }
      P.steps.push_back(s);
    }
  };

  int wacc_next = 0;
  (void)merge_k;
  const bool fused_latent = M == 1 && Z <= 16;     // head -> latent and d/dz -> latent backward stay in registers
  const int nl = a.non_linear;
  (void)nl;

  // ================= forward =================
  for (int m = 0; m < M; ++m) {
    const ModDesc& q = a.mod[m];
    for (int l = 0; l < L; ++l) {
      for (int h = 0; h < 2; ++h) {
        if (l == 0) emit_fwd(h, w_enc[m][0], SP_X, 0, 0, m);
        else emit_fwd(h, w_enc[m][l], SP_NONE, 0, 0, 0);
        Epi e = new_epi(EK_HIDDEN, h, accbuf(h), m);
        e.n_mma = w_enc[m][l].R; e.n_valid = q.enc[l].out; e.n_cols = round16(q.enc[l].out + 1);
        e.to_act = 1; e.stash_off = fwd_only ? -1 : s_h[m][l * 2 + h];
        e.src_cg = is64(q.enc[l].out + 1) ? e.n_cols / 8 : 0;
        const int id = push_epi(e);
        act_ready[h] = id; acc_free[accbuf(h)] = id;
      }
    }
    for (int h = 0; h < 2 && fused_latent; ++h) {
      emit_fwd(h, w_head[m], SP_NONE, 0, 0, 0);
      Epi e = new_epi(EK_HEAD_LATENT, h, accbuf(h), m);
      e.n_mma = w_head[m].R; e.to_act = 1;
      const int id = push_epi(e);
      act_ready[h] = id; acc_free[accbuf(h)] = id;
      ready[s_g0[m * 2 + h]] = id;
    }
    for (int h = 0; h < 2 && !fused_latent; ++h) {
      emit_fwd(h, w_head[m], SP_NONE, 0, 0, 0);
      Epi e = new_epi(EK_HEAD, h, accbuf(h), m);
      e.n_mma = w_head[m].R; e.n_valid = 2 * Z; e.n_cols = round16(2 * Z);
      const int id = push_epi(e);
      acc_free[accbuf(h)] = id;
      if (act_ready[h] < id) act_ready[h] = id;      // ACT[h] may be overwritten only after the head read it
    }
  }
  for (int h = 0; h < 2 && !fused_latent; ++h) {
    Epi e = new_epi(EK_LATENT, h, -1, 0);
    e.to_act = 1;
    const int id = push_epi(e);
    act_ready[h] = id;
    for (int m = 0; m < M; ++m) ready[s_g0[m * 2 + h]] = id;
  }
  std::vector<int> dxh_ready(M * 2, 0);   // [m*2+h]: item after which all dxh blocks of (m,h) exist
  for (int m = 0; m < M; ++m) {
    const ModDesc& q = a.mod[m];
    if (m > 0) {
      for (int h = 0; h < 2; ++h) {
        Epi e = new_epi(EK_COPY, h, -1, m);
        e.src_off = s_g0[m * 2 + h]; e.src_cg = round16(Z + C + 1) / 8;
        act_ready[h] = push_epi(e);
      }
    }
    for (int l = 0; l < L; ++l) {
      for (int h = 0; h < 2; ++h) {
        emit_fwd(h, w_dec[m][l], SP_NONE, 0, 0, 0);
        Epi e = new_epi(EK_HIDDEN, h, accbuf(h), m);
        e.n_mma = w_dec[m][l].R; e.n_valid = q.dec[l].out; e.n_cols = round16(q.dec[l].out + 1);
        e.to_act = 1; e.stash_off = fwd_only ? -1 : s_k[m][l * 2 + h];
        e.src_cg = is64(q.dec[l].out + 1) ? e.n_cols / 8 : 0;
        const int id = push_epi(e);
        act_ready[h] = id; acc_free[accbuf(h)] = id;
      }
    }
    if (lay.n_dxh_blk[m] == 0) {          // fast path: one tile, d/dx_recon planes land in ACT[h]
      for (int h = 0; h < 2; ++h) {
        emit_fwd(h, w_out[m][0], SP_NONE, 0, 0, 0);
        Epi e = new_epi(fwd_only ? EK_XHAT : EK_RECON, h, accbuf(h), m);
        e.n_mma = w_out[m][0].R; e.n_valid = q.D; e.n_cols = round16(q.D); e.col0 = 0;
        e.to_act = fwd_only ? 0 : 1; e.last = 1;
        e.src_cg = (M == 1 && !fwd_only) ? 1 : 0;   // one modality: last forward item of the half, it publishes the stash early
        const int id = push_epi(e);
        act_ready[h] = id; acc_free[accbuf(h)] = id;
      }
    }
    // Wide output layer: x_recon in tiles of 64 columns.  Consecutive tiles of a half alternate between the two 64-column
    // halves of acc[h] (each with its own barrier), so the MMAs of tile t + 1 run while the epilogue of tile t -- the
    // longest item of the forward pass -- is still reading tile t.
    int sub_free[2][2] = {{acc_free[0], acc_free[0]}, {acc_free[1], acc_free[1]}};
    for (int t = 0; t < lay.n_dxh_blk[m]; ++t) {
      for (int h = 0; h < 2; ++h) {
        const int sb = split_recon ? (t & 1) : 0;
        const int bufid = sb ? 4 + h : accbuf(h);
        {   // x_recon columns 64 t .. 64 t + 63: ACT[h] (K-major) x 8 column groups of the transposed block (MN-major)
          const WRef& w = w_out[m][0];
          Step s = base_step(h);
          set_b(s, SP_W, w.wp_off, w.R, 8 * t, 8, true);
          set_a_kmajor(s, 0);
          s.ksteps = (unsigned short)(w.R / 16); s.n = 64; s.tmem_col = (unsigned short)(kAcc0 + 128 * h + 64 * sb);
          s.first = 1;
          need(s, sub_free[h][sb]); need(s, act_ready[h]);
          s.commit = 1; s.commit_buf = (unsigned char)bufid;
          P.steps.push_back(s);
        }
        Epi e = new_epi(fwd_only ? EK_XHAT : EK_RECON, h, bufid, m);
        e.n_mma = 64; e.n_valid = q.D - 64 * t < 64 ? q.D - 64 * t : 64; e.n_cols = 64; e.col0 = 64 * t;
        e.stash_off = fwd_only ? -1 : lay.dxh_blk[m] + ((long long)h * lay.n_dxh_blk[m] + t) * 32768;
        e.last = t == lay.n_dxh_blk[m] - 1;
        const int id = push_epi(e);
        acc_free[accbuf(h)] = id;
        sub_free[h][sb] = id;
        if (!split_recon) sub_free[h][1] = id;
        dxh_ready[m * 2 + h] = id;
      }
    }
  }

  if (fwd_only) {
    push_epi(new_epi(EK_STEP_END, 2, -1, 0));
    for (Step& s : P.steps) s.dep_grp = 2;
    return P;
  }
  // ================= backward =================
  // Generic-proxy stores to the stash become visible to the TMA (async proxy) at the per-half EK_FENCE item:
  // every stash-sourced tile of half h waits for it (all of them are consumed in the backward pass).
  int fence_id[2];
  for (int h = 0; h < 2; ++h) {
    Epi e = new_epi(EK_FENCE, h, -1, 0);
    e.src_cg = (M == 1 && lay.n_dxh_blk[0] == 0) ? 1 : 0;      // already published by the reconstruction item
    fence_id[h] = push_epi(e);
  }
  // logvar_out: its gradient partials are complete once both halves have finished their reconstruction items.  The
  // optimiser group, idle until the first weight gradient arrives, applies Adam to it here instead of inside the
  // step-end rendezvous (n_valid / n_cols: the items of groups 0 / 1 it waits for).
  if (a.loss_kind == NMB_LOSS_GAUSS_LL)
    for (int m = 0; m < M; ++m) {
      Epi e = new_epi(EK_LAM, 2, -1, m);
      e.n_valid = fence_id[0]; e.n_cols = fence_id[1];
      push_epi(e);
    }

  // One linear layer: [dgrad(h) ->] wgrad parts(h) per half, then the Adam items.  The dgrad MMAs come first
  // so that they have consumed the pre-update weight planes before the wgrad accumulator is committed (the
  // Adam epilogue waits for that commit); acc[h] is committed after the wgrad part, so the epilogue that
  // overwrites ACT[h] cannot start before the wgrad MMAs have read it.
  //   dg_kind: 0 = no data gradient (first encoder layer), 1 = EK_DGRAD, 2 = EK_DZ (decoder input: d/dz only)
  //   guard_act: no dgrad, but an epilogue-only item of the same half overwrites ACT[h] next (EK_COPY of the next
  //   modality): the wgrad part commits acc[h] and that item waits for it.
  auto layer_backward = [&](int m, const LinDesc& w, const WRef& wr, int b_space, const long long in_base[2],
                            int dg_kind, int n_need, int x_mod, bool guard_act, bool blk64) {
    const int in_cg = round16(w.in + 1) / 8;
    const int b64 = blk64 ? in_cg : 0;
    const int n_items = (in_cg + 15) / 16;
    for (int it0 = 0; it0 < n_items; it0 += 2) {
      const int it1 = it0 + 2 < n_items ? it0 + 2 : n_items;
      const bool last_pair = it1 == n_items;
      const int wb0 = wacc_next;
      for (int h = 0; h < 2; ++h) {
        if (dg_kind && it0 == 0) emit_dgrad(h, wr, n_need, false);
        int wb = wb0;
        for (int it = it0; it < it1; ++it, wb ^= 1) {
          const int gf = it * 16, gc = in_cg - gf < 16 ? in_cg - gf : 16;
          emit_wgrad_part(h, wb, b_space, in_base[h], gf, gc, x_mod, (dg_kind || guard_act) && last_pair && it == it1 - 1, b64);
        }
        if (dg_kind && last_pair) {
          Epi e = new_epi(dg_kind == 1 ? EK_DGRAD : (fused_latent ? EK_DZ_LATENT_BWD : EK_DZ), h, accbuf(h), m);
          e.n_mma = round16(n_need); e.n_valid = n_need;
          if (dg_kind == 1) { e.n_cols = round16(n_need); e.to_act = 1; e.src_off = in_base[h]; e.src_cg = blk64 ? in_cg : 0; }
          else if (fused_latent) e.to_act = 1;
          const int id = push_epi(e);
          act_ready[h] = id; acc_free[accbuf(h)] = id;
        }
      }
      int wb = wb0;
      for (int it = it0; it < it1; ++it, wb ^= 1) {
        Epi e = new_epi(EK_WGRAD, 2, 2 + wb, m);
        e.col0 = it * 128; e.n_mma = (in_cg - it * 16 < 16 ? in_cg - it * 16 : 16) * 8;
        e.p_off = w.off; e.p_ld = w.ld; e.p_rows = w.out; e.p_cols = w.in + 1;
        e.wp_off = wr.wp_off; e.wp_R = wr.R;
        e.mst_off = mst[w.off].mst_off; e.mst_R = mst[w.off].R;
        acc_free[2 + wb] = push_epi(e);
      }
      if ((it1 - it0) & 1) wacc_next ^= 1;
    }
  };

  for (int m = 0; m < M; ++m) {
    const ModDesc& q = a.mod[m];
    const long long k_last[2] = {s_k[m][(L - 1) * 2 + 0], s_k[m][(L - 1) * 2 + 1]};
    if (lay.n_dxh_blk[m] == 0) {
      // decoder_mean_layer as a regular layer: ACT[h] holds d/dx_recon (written by EK_RECON)
      layer_backward(m, q.outl, w_out[m][0], SP_STASH, k_last, 1, q.outl.in, 0, false, is64(q.outl.in + 1));
    } else {
      if (M > 1) {   // ACT[h] <- last decoder hidden activation of this modality (transposed wgrad A operand)
        for (int h = 0; h < 2; ++h) {
          Epi e = new_epi(EK_COPY, h, -1, m);
          e.src_off = k_last[h]; e.src_cg = round16(q.outl.in + 1) / 8;
          act_ready[h] = push_epi(e);
        }
      }
      // Per 64-column block t of d/dx_recon (one 32 KB stash tile) and half h: the data-gradient MMAs (tile = K-major
      // A, weights of block t = MN-major B, accumulated over t in acc[h]) and the transposed weight-gradient MMAs
      // (ACT[h]^T x the SAME tile as MN-major B) run back to back on one load of the tile: the first step keeps the
      // tile's ring slot (a_hold), the second reads it (b_held) and releases it.  Blocks go in pairs (one weight-gradient
      // accumulator = 128 output rows); every data-gradient MMA of a pair is issued before the pair's Adam item can
      // start, so the weight planes it rewrites have been consumed.  acc[h] is committed by the last transposed part of
      // the half: the EK_DGRAD epilogues, which overwrite ACT[h], come last.
      for (int t0 = 0; t0 < lay.n_dxh_blk[m]; t0 += 2) {
        const int wb = wacc_next; wacc_next ^= 1;
        const int nt = lay.n_dxh_blk[m] - t0 < 2 ? lay.n_dxh_blk[m] - t0 : 2;
        const bool last_item = t0 + 2 >= lay.n_dxh_blk[m];
        for (int h = 0; h < 2; ++h) {
          const int buf = accbuf(h);
          for (int t = t0; t < t0 + nt; ++t) {
            const WRef& w = w_out[m][0];
            const long long tile = lay.dxh_blk[m] + ((long long)h * lay.n_dxh_blk[m] + t) * 32768;
            Step s = base_step(h);
            set_b(s, SP_W, w.wp_off, w.R, 8 * t, 8, false);       // rows = input index (N), 8 groups = 64 output rows (K)
            set_a_kmajor(s, 0);
            s.a_space = SP_STASH; s.a_off = tile; s.a_bytes = 32768; s.a_hold = 1;
            s.ksteps = 4; s.n = (unsigned short)w.R; s.tmem_col = (unsigned short)(kAcc0 + 128 * buf);
            s.first = t == 0;
            if (t == 0) need(s, acc_free[buf]);
            s.commit_buf = (unsigned char)buf;
            P.steps.push_back(s);

            Step g = base_step(h);
            set_b(g, SP_NONE, 0, 128, 0, 8, true);
            g.b_bytes = 0; g.b_held = 1;
            set_a_mnmajor(g);
            g.ksteps = 8; g.n = 64; g.tmem_col = (unsigned short)(kWacc0 + 128 * wb + 64 * (t - t0));
            g.first = h == 0;
            if (t == t0) { need(g, act_ready[h]); need(g, acc_free[2 + wb]); }
            const bool last = t == t0 + nt - 1;
            g.commit = !last ? 0 : (h == 1 ? 1 : 2);
            g.commit_buf = (unsigned char)(2 + wb);
            if (last && last_item) { g.commit2 = 1; g.commit2_buf = (unsigned char)accbuf(h); }
            P.steps.push_back(g);
          }
        }
        Epi e = new_epi(EK_WGRAD_T, 2, 2 + wb, m);
        e.n_mma = 64 * nt; e.col0 = 64 * t0;
        e.p_off = q.outl.off; e.p_ld = q.outl.ld; e.p_rows = q.outl.out; e.p_cols = q.outl.in + 1;
        e.wp_off = w_out[m][0].wp_off; e.wp_R = w_out[m][0].R;
        e.mst_off = mst[q.outl.off].mst_off; e.mst_R = mst[q.outl.off].R;
        acc_free[2 + wb] = push_epi(e);
      }
      for (int h = 0; h < 2; ++h) {
        const int buf = accbuf(h);
        Epi e = new_epi(EK_DGRAD, h, buf, m);
        e.n_mma = round16(q.outl.in + 1); e.n_valid = q.outl.in; e.n_cols = round16(q.outl.in);
        e.to_act = 1; e.src_off = k_last[h];
        e.src_cg = is64(q.outl.in + 1) ? round16(q.outl.in + 1) / 8 : 0;
        const int id = push_epi(e);
        act_ready[h] = id; acc_free[buf] = id;
      }
    }
    // decoder hidden layers
    for (int l = L - 1; l >= 0; --l) {
      const long long in_base[2] = {l == 0 ? s_g0[m * 2 + 0] : s_k[m][(l - 1) * 2 + 0],
                                    l == 0 ? s_g0[m * 2 + 1] : s_k[m][(l - 1) * 2 + 1]};
      layer_backward(m, q.dec[l], w_dec[m][l], SP_STASH, in_base, l > 0 ? 1 : 2, l > 0 ? q.dec[l].in : Z, 0, false,
                     l > 0 && is64(q.dec[l].in + 1));
    }
  }
  for (int h = 0; h < 2 && !fused_latent; ++h) {
    Epi e = new_epi(EK_LATENT_BWD, h, -1, 0);
    e.to_act = 1;
    act_ready[h] = push_epi(e);
  }
  for (int m = 0; m < M; ++m) {
    const ModDesc& q = a.mod[m];
    if (m > 0) {
      for (int h = 0; h < 2; ++h) {
        Epi e = new_epi(EK_COPY, h, accbuf(h), m);      // waits for the previous modality's last wgrad MMAs (guard_act)
        e.src_off = s_dmulv[m * 2 + h]; e.src_cg = round16(2 * Z) / 8;
        const int id = push_epi(e);
        act_ready[h] = id; acc_free[accbuf(h)] = id;
      }
    }
    // head, then encoder hidden layers
    for (int l = L; l >= 0; --l) {
      const bool is_head = l == L;
      const LinDesc& w = is_head ? q.head : q.enc[l];
      const WRef& wr = is_head ? w_head[m] : w_enc[m][l];
      if (!is_head && l == 0) {
        const long long xb[2] = {0, 0};
        layer_backward(m, w, wr, SP_X, xb, 0, 0, m, m + 1 < M, false);
      } else {
        const int li = (is_head ? L : l) - 1;
        const long long in_base[2] = {s_h[m][li * 2 + 0], s_h[m][li * 2 + 1]};
        layer_backward(m, w, wr, SP_STASH, in_base, 1, w.in, 0, false, is64(w.in + 1));
      }
    }
  }
  // logvar_out (needs the partial sums of both halves) is updated inside EK_STEP_END, after the rendezvous
  {   // Adam items every group helps with (column split): those behind the last per-half item of the step (both
      // activation groups are idle by then) and the transposed weight gradients of decoder_mean_layer (the
      // activation groups' next items wait on exactly those MMAs and Adam items).  A group joining a shared item
      // first waits until the optimiser has finished its last optimiser-only item before it (wait_optim), so that
      // its view of the accumulator barrier phases cannot alias.
    int last_half_item = -1;
    for (int k = 0; k < (int)P.epis.size(); ++k) if (P.epis[k].half != 2) last_half_item = k;
    int last_optim_only = 0;
    for (int k = 0; k < (int)P.epis.size(); ++k) {
      Epi& ep = P.epis[k];
      if (ep.kind != EK_WGRAD && ep.kind != EK_WGRAD_T) continue;
      if (k > last_half_item || ep.kind == EK_WGRAD_T) { ep.split_all = 1; ep.wait_optim = last_optim_only; }
      else last_optim_only = k + 1;
    }
  }
  // an accumulator consumed by a shared (split_all) Adam item is free only when ALL groups are done with it:
  // encoded as a negative optimiser dependency
  for (Step& st : P.steps)
    if (st.mma_dep_joint > 0 && P.epis[st.mma_dep_joint - 1].split_all) st.mma_dep_joint = -st.mma_dep_joint;
  push_epi(new_epi(EK_STEP_END, 2, -1, 0));
  for (Step& s : P.steps) {
    s.dep_grp = 2;
    if (s.a_space == SP_STASH || s.b_space == SP_STASH) { s.dep = fence_id[s.half]; s.dep_grp = s.half; }
  }
  return P;
}

}  // namespace tcp
}  // namespace nmb
