#include "nmb_tcp.h"
#include <cstdio>
#include <cstdlib>
using namespace nmb;
int main(int argc, char** argv){
  NmbArch a{}; a.n_mod = argc > 2 ? atoi(argv[2]) : 1;
  for (int m = 0; m < a.n_mod; ++m) a.input_dims[m] = atoi(argv[1]);
  a.n_hidden=2; a.hidden[0]=110; a.hidden[1]=110; a.latent=10; a.c_dim=29; a.non_linear=1; a.combine = 1;
  ArchDesc d; const char* err; if(build_arch(a,&d,&err)){printf("err %s\n",err);return 1;}
  auto P = tcp::build_program(d);
  const int NS = P.steps.size(), NE = P.epis.size();
  printf("steps %d epis %d\n", NS, NE);
  // tile list
  struct Tile { int step; int dep; int grp; bool w; };
  std::vector<Tile> tiles; std::vector<int> first_tile(NS), n_tiles(NS);
  for (int k = 0; k < NS; ++k) { const auto& s = P.steps[k]; first_tile[k] = tiles.size();
    if (s.a_bytes) tiles.push_back({k, s.dep, s.dep_grp, false});
    tiles.push_back({k, s.dep, s.dep_grp, s.b_space == tcp::SP_W}); n_tiles[k] = tiles.size() - first_tile[k]; }
  int issued = 0, consumed = 0, mi = 0; int epi_done[3] = {0, 0, 0}; int commits[4] = {0,0,0,0}; int waited[3][4] = {{0}};
  int ep[3] = {0, 0, 0};
  auto owner = [&](const tcp::Epi& e, int g) {
    if (e.kind == tcp::EK_STEP_END) return g;
    if (e.kind == tcp::EK_WGRAD || e.kind == tcp::EK_WGRAD_T) return 2;
    return e.half;
  };
  for (int iter = 0; iter < 100000; ++iter) {
    bool prog = false;
    // producer
    while (issued < (int)tiles.size() && issued - consumed < 3) {
      const Tile& t = tiles[issued]; bool ok = true;
      if (t.dep) { if (t.grp < 2) ok = epi_done[t.grp] >= t.dep; else ok = epi_done[0] >= t.dep && epi_done[1] >= t.dep && epi_done[2] >= t.dep; }
      if (!ok) break; ++issued; prog = true; }
    // mma
    while (mi < NS) { const auto& s = P.steps[mi];
      if (s.mma_dep && epi_done[s.half] < s.mma_dep) break;
      if (s.mma_dep_joint && epi_done[2] < s.mma_dep_joint) break;
      if (issued < first_tile[mi] + n_tiles[mi]) break;
      consumed += n_tiles[mi]; if (s.commit == 1) commits[s.commit_buf]++; if (s.commit2) commits[s.half]++; ++mi; prog = true; }
    // epilogue groups
    for (int g = 0; g < 3; ++g) {
      while (ep[g] < NE) { const auto& e = P.epis[ep[g]];
        if (owner(e, g) != g) { ++ep[g]; continue; }
        if (e.kind == tcp::EK_STEP_END) {        // rendezvous: every group must be waiting at this item
          bool all = true;
          for (int o = 0; o < 3; ++o) {
            int q = ep[o]; while (q < NE && owner(P.epis[q], o) != o) ++q;
            if (q != ep[g]) all = false;
          }
          if (!all) break;
          for (int o = 0; o < 3; ++o) { epi_done[o] = ep[g] + 1; ep[o] = ep[g] + 1; }   // released together
          prog = true; continue;
        }
        if (e.buf >= 0) { if (commits[e.buf] <= waited[g][e.buf]) break; waited[g][e.buf]++; }
        epi_done[g] = ep[g] + 1; ++ep[g]; prog = true; }
    }
    if (mi == NS && ep[0] == NE && ep[1] == NE && ep[2] == NE) { printf("OK all done (iters %d)\n", iter); return 0; }
    if (!prog) { printf("STUCK: issued %d consumed %d mma step %d (dep %d optim-dep %d half %d) ep %d %d %d epi_done %d %d %d commits %d %d %d %d\n",
       issued, consumed, mi, mi < NS ? P.steps[mi].mma_dep : -1, mi < NS ? P.steps[mi].mma_dep_joint : -1, mi < NS ? P.steps[mi].half : -1,
       ep[0], ep[1], ep[2], epi_done[0], epi_done[1], epi_done[2], commits[0], commits[1], commits[2], commits[3]); return 1; }
  }
  return 0;
}
