#include "nmb_tcp.h"
#include <cstdio>
#include <cstdlib>
using namespace nmb;
int main(int argc, char** argv){
  NmbArch a{}; a.n_mod = argc > 2 ? atoi(argv[2]) : 1;
  for (int m = 0; m < a.n_mod; ++m) a.input_dims[m] = atoi(argv[1]);
  a.n_hidden=2; a.hidden[0]=110; a.hidden[1]=110; a.latent=10; a.c_dim=29; a.non_linear=1; a.combine = 1;
  ArchDesc d; const char* err; if(build_arch(a,&d,&err)){printf("err %s\n",err);return 1;}
  const bool fwd = argc > 3 && argv[3][0] == 'f';      // forward-only (reconstruction) program
  auto P = tcp::build_program(d, fwd);
  const int NS = P.steps.size(), NE = P.epis.size();
  printf("steps %d epis %d\n", NS, NE);
  // the compact item table that travels in the kernel parameters must reproduce every field the kernel reads
  for (int k = 0; k < NE; ++k) {
    const tcp::Epi& e = P.epis[k];
    if (!tcp::epip_fits(e)) { printf("item %d does not fit the compact form\n", k); return 5; }
    const tcp::Epi r = tcp::from_epip(tcp::to_epip(e));
    const bool same = r.kind == e.kind && r.half == e.half && r.buf == e.buf && r.mod == e.mod && r.n_mma == e.n_mma &&
                      r.n_valid == e.n_valid && r.n_cols == e.n_cols && r.tmem_col == e.tmem_col && r.col0 == e.col0 &&
                      r.to_act == e.to_act && r.stash_off == e.stash_off && r.src_off == e.src_off && r.src_cg == e.src_cg &&
                      r.p_off == e.p_off && r.p_ld == e.p_ld && r.p_rows == e.p_rows && r.p_cols == e.p_cols &&
                      r.wp_off == e.wp_off && r.wp_R == e.wp_R && r.mst_off == e.mst_off && r.mst_R == e.mst_R &&
                      r.row0 == e.row0 && r.last == e.last && r.split_all == e.split_all && r.wait_optim == e.wait_optim;
    if (!same) { printf("item %d: compact form loses a field\n", k); return 6; }
  }
  // a step that reads the tile kept by its predecessor must directly follow a step that keeps one, in the same half
  for (int k = 0; k < NS; ++k)
    if (P.steps[k].b_held && (k == 0 || !P.steps[k - 1].a_hold || P.steps[k - 1].half != P.steps[k].half)) {
      printf("step %d reads a kept tile nobody kept\n", k); return 7; }
  // tile list
  struct Tile { int step; int dep; int grp; bool w; };
  std::vector<Tile> tiles; std::vector<int> first_tile(NS), n_tiles(NS);
  for (int k = 0; k < NS; ++k) { const auto& s = P.steps[k]; first_tile[k] = tiles.size();
    if (s.a_bytes) tiles.push_back({k, s.dep, s.dep_grp, false});
    if (s.b_bytes) tiles.push_back({k, s.dep, s.dep_grp, s.b_space == tcp::SP_W}); n_tiles[k] = tiles.size() - first_tile[k]; }
  auto owner = [&](const tcp::Epi& e, int g) {
    if (e.kind == tcp::EK_STEP_END || e.split_all) return g;
    if (e.kind == tcp::EK_WGRAD || e.kind == tcp::EK_WGRAD_T) return 2;
    return e.half;
  };
  // accumulator hazard check: the MMAs of use #u of an accumulator buffer may only be issued when every group that
  // owns the epilogue item of use #u-1 has finished reading it
  std::vector<int> uses[6];
  for (int k = 0; k < NE; ++k) if (P.epis[k].buf >= 0) uses[P.epis[k].buf].push_back(k);
  // Randomised interleaving: each trial advances one randomly chosen role by one unit at a time, so that any
  // ordering the hardware could produce between the roles is sampled (300 seeds).
  for (unsigned seed = 1; seed <= 300; ++seed) {
    unsigned rng = seed * 2654435761u;
    auto rnd = [&]() { rng ^= rng << 13; rng ^= rng >> 17; rng ^= rng << 5; return rng; };
    int issued = 0, consumed = 0, mi = 0; int epi_done[3] = {0, 0, 0}; int commits[6] = {0,0,0,0,0,0}; int waited[3][6] = {{0}};
    int ep[3] = {0, 0, 0}; int use_idx[6] = {0, 0, 0, 0, 0, 0};
    auto step_producer = [&]() {
      if (!(issued < (int)tiles.size() && issued - consumed < 3)) return false;
      const Tile& t = tiles[issued];
      if (t.dep) { const bool ok = t.grp < 2 ? epi_done[t.grp] >= t.dep : (epi_done[0] >= t.dep && epi_done[1] >= t.dep && epi_done[2] >= t.dep); if (!ok) return false; }
      ++issued; return true; };
    auto step_mma = [&]() -> int {
      if (mi >= NS) return 0;
      const auto& s = P.steps[mi];
      if (s.mma_dep && epi_done[s.half] < s.mma_dep) return 0;
      if (s.mma_dep_joint > 0 && epi_done[2] < s.mma_dep_joint) return 0;
      if (s.mma_dep_joint < 0 && (epi_done[0] < -s.mma_dep_joint || epi_done[1] < -s.mma_dep_joint || epi_done[2] < -s.mma_dep_joint)) return 0;
      if (issued < first_tile[mi] + n_tiles[mi]) return 0;
      // accumulator buffer id: the committing step names it (4 / 5 = upper halves of acc[h]); partial steps by column
      const int b = s.commit == 1 ? s.commit_buf : s.tmem_col / 128, u = use_idx[b];
      if (u > 0) {
        const int prev = uses[b][u - 1];
        for (int g = 0; g < 3; ++g)
          if (owner(P.epis[prev], g) == g && ep[g] <= prev) { printf("HAZARD (seed %u): step %d overwrites accumulator %d before group %d finished item %d\n", seed, mi, b, g, prev); return -1; }
      }
      consumed += n_tiles[mi] - (s.a_hold ? 1 : 0) + (s.b_held ? 1 : 0);      // a kept slot is released by the next step
      if (s.commit == 1) { commits[s.commit_buf]++; use_idx[s.commit_buf]++; }
      if (s.commit2) { commits[s.half]++; use_idx[s.half]++; }
      ++mi; return 1; };
    auto step_group = [&](int g) {
      while (ep[g] < NE && owner(P.epis[ep[g]], g) != g) {
        const auto& e = P.epis[ep[g]];
        if (e.kind == tcp::EK_WGRAD || e.kind == tcp::EK_WGRAD_T) waited[g][e.buf]++;
        ++ep[g];
      }
      if (ep[g] >= NE) return false;
      const auto& e = P.epis[ep[g]];
      if (e.kind == tcp::EK_STEP_END) {
        for (int o = 0; o < 3; ++o) {
          int q = ep[o]; while (q < NE && owner(P.epis[q], o) != o) ++q;
          if (q != ep[g]) return false;
        }
        for (int o = 0; o < 3; ++o) { epi_done[o] = ep[g] + 1; ep[o] = ep[g] + 1; }
        return true;
      }
      if (e.split_all && g != 2 && e.wait_optim && epi_done[2] < e.wait_optim) return false;
      if (e.kind == tcp::EK_LAM && (epi_done[0] < e.n_valid || epi_done[1] < e.n_cols)) return false;   // waits for both halves
      if (e.buf >= 0) {
        if (commits[e.buf] <= waited[g][e.buf]) return false;
        if (commits[e.buf] - waited[g][e.buf] > 1) { printf("PHASE ALIAS (seed %u): group %d item %d\n", seed, g, ep[g]); exit(4); }
        waited[g][e.buf]++;
      }
      epi_done[g] = ep[g] + 1; ++ep[g]; return true; };
    int idle = 0;
    while (!(mi == NS && ep[0] >= NE && ep[1] >= NE && ep[2] >= NE)) {
      const unsigned r = rnd() % 5;
      int ok = 0;
      if (r == 0) ok = step_producer(); else if (r == 1) { ok = step_mma(); if (ok < 0) return 3; } else ok = step_group((int)r - 2);
      idle = ok ? 0 : idle + 1;
      if (idle > 200) {
        const bool any = step_producer() || step_mma() > 0 || step_group(0) || step_group(1) || step_group(2);
        if (!any) { printf("STUCK (seed %u): issued %d consumed %d mma step %d ep %d %d %d epi_done %d %d %d\n", seed, issued, consumed, mi, ep[0], ep[1], ep[2], epi_done[0], epi_done[1], epi_done[2]); return 1; }
        idle = 0;
      }
    }
  }
  printf("OK all done (300 random interleavings)\n");
  return 0;
}
