/* cuda_emul.cpp - TEST INFRASTRUCTURE ONLY: fibre scheduler behind cuda_emul.h. */
#include "cuda_emul.h"

#include <ucontext.h>
#include <vector>

namespace emu {

uint3 g_tid, g_bid;
dim3 g_bdim, g_gdim;
unsigned char* g_dyn_smem = nullptr;

namespace {
constexpr size_t kStack = 256 * 1024;

struct Fibre {
  ucontext_t ctx;
  unsigned char* stack = nullptr;
  bool done = false;
  uint3 tid;
};

std::vector<Fibre> fibres;
ucontext_t sched_ctx;
int cur = -1;
const std::function<void()>* body_ptr = nullptr;

/* CTA barrier */
unsigned long bar_gen = 0;
int bar_arrived = 0, live = 0;

/* per-warp rendezvous */
struct WarpState {
  unsigned long gen = 0;
  int arrived = 0;
  int live = 0;
  unsigned long long slot[32];
  unsigned long long pub[32];
  unsigned ballot_acc = 0, ballot_pub = 0;
};
std::vector<WarpState> warps;

void yield() { swapcontext(&fibres[cur].ctx, &sched_ctx); }

void trampoline() {
  (*body_ptr)();
  Fibre& f = fibres[cur];
  f.done = true;
  live--;
  warps[cur / 32].live--;
  /* a thread that exits must not leave others stuck at a barrier it will never reach */
  if (bar_arrived > 0 && bar_arrived >= live) { bar_arrived = 0; bar_gen++; }
  WarpState& w = warps[cur / 32];
  if (w.arrived > 0 && w.arrived >= w.live) {
    memcpy(w.pub, w.slot, sizeof(w.pub)); w.ballot_pub = w.ballot_acc; w.ballot_acc = 0;
    w.arrived = 0; w.gen++;
  }
  swapcontext(&f.ctx, &sched_ctx);
}

void warp_rendezvous() {
  WarpState& w = warps[cur / 32];
  unsigned long my = w.gen;
  if (++w.arrived >= w.live) {
    memcpy(w.pub, w.slot, sizeof(w.pub));
    w.ballot_pub = w.ballot_acc; w.ballot_acc = 0;
    w.arrived = 0; w.gen++;
    return;
  }
  while (w.gen == my) yield();
}
}  // namespace

void sync_threads() {
  unsigned long my = bar_gen;
  if (++bar_arrived >= live) { bar_arrived = 0; bar_gen++; return; }
  while (bar_gen == my) yield();
}

void sync_warp() { warp_rendezvous(); }

unsigned long long warp_exchange(unsigned long long v, int src_lane) {
  WarpState& w = warps[cur / 32];
  w.slot[cur & 31] = v;
  warp_rendezvous();
  unsigned long long r = w.pub[src_lane & 31];
  return r;
}

unsigned ballot(int pred) {
  WarpState& w = warps[cur / 32];
  if (pred) w.ballot_acc |= 1u << (cur & 31);
  warp_rendezvous();
  return w.ballot_pub;
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const unsigned nthreads = block.x * block.y * block.z;
  if (fibres.size() < nthreads) {
    size_t old = fibres.size();
    fibres.resize(nthreads);
    for (size_t i = old; i < nthreads; i++) fibres[i].stack = (unsigned char*)malloc(kStack);
  }
  warps.assign((nthreads + 31) / 32, WarpState());
  unsigned char* dyn = (unsigned char*)calloc(smem ? smem : 1, 1);
  g_dyn_smem = dyn;
  g_bdim = block; g_gdim = grid;
  body_ptr = &body;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        g_bid = {bx, by, bz};
        live = (int)nthreads; bar_arrived = 0;
        for (auto& w : warps) { w = WarpState(); }
        for (unsigned t = 0; t < nthreads; t++) {
          Fibre& f = fibres[t];
          f.done = false;
          f.tid = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
          warps[t / 32].live++;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        while (live > 0) {
          for (unsigned t = 0; t < nthreads; t++) {
            if (fibres[t].done) continue;
            cur = (int)t;
            g_tid = fibres[t].tid;
            swapcontext(&sched_ctx, &fibres[t].ctx);
          }
        }
      }
  cur = -1;
  g_dyn_smem = nullptr;
  free(dyn);
}

}  // namespace emu
