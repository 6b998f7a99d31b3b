// Parity of the multi-GPU range policy RAJA::b200_multi_exec (include/RAJA_b200/multi.hpp) against the reference's seq_exec
// run over the WHOLE (unsharded) arrays on the host.  Ranks are host threads of this process, one per visible device
// (RB200_WORLD overrides; with one GPU the world is 1 and every exchange path still runs, against the rank itself):
//   * forall over the shard with global indices; expt::Reduce sum / min / max / ReduceLoc across ranks (bit-exact for integers,
//     n eps for fp64; first-index tie rule across rank boundaries)
//   * inclusive / exclusive scans with the carry exchanged between ranks (int plus, int minimum, double plus)
//   * multi_sort_pairs (u32 keys with heavy duplication + global index as value: stability across ranks), multi_sort fp64 descending
//   * shards of different lengths, an empty shard, a world that does not divide N
#include "RAJA_b200/RAJA.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <unistd.h>
#include <vector>

static std::mutex g_mu;
static int g_fail = 0;
#define EXPECT(cond)                                                                                  \
  do {                                                                                                \
    if (!(cond)) { std::lock_guard<std::mutex> l(g_mu); printf("FAIL rank %d %s:%d: %s\n", rank, __FILE__, __LINE__, #cond); ++g_fail; } \
  } while (0)
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// Every copy of a rank goes through the rank's OWN stream: the synchronous cudaMemcpy / cudaMemset run on the legacy default
// stream, which waits for all blocking streams of the device -- with several ranks on one device (RB200_OVERSUBSCRIBE) that is
// another rank's exchange kernel, spinning until THIS rank arrives.
static thread_local cudaStream_t tl_stream = nullptr;
static void copy_sync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind)
{
  if (!bytes) return;
  CK(cudaMemcpyAsync(dst, src, bytes, kind, tl_stream));
  CK(cudaStreamSynchronize(tl_stream));
}
template <class T> T* to_dev(const T* h, size_t n) { T* p; CK(cudaMalloc(&p, n * sizeof(T) + 64)); copy_sync(p, h, n * sizeof(T), cudaMemcpyHostToDevice); return p; }
template <class T> std::vector<T> to_host(const T* d, size_t n) { std::vector<T> h(n); copy_sync(h.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost); return h; }

using MEXEC = RAJA::b200_multi_exec<256>;

// pieces of every rank, collected for the checks of rank 0
struct Collect {
  std::vector<std::vector<unsigned>> k32, v32;
  std::vector<std::vector<double>> f64;
};

static void rank_main(int rank, int world, RAJA::b200::LocalPeerGroup* group, Collect* col, long N)
{
  CK(cudaSetDevice(rank % [] { int n = 1; cudaGetDeviceCount(&n); return n; }()));
  void* const* ptrs = group->attach(rank);
  // camp's default Cuda resource draws from a stream pool created on whichever device was current first: a rank that owns
  // another device brings its own stream (camp::resources::Cuda::CudaFromStream, tpl/camp/include/camp/resource/cuda.hpp:140-146)
  cudaStream_t stream;
  CK(cudaStreamCreate(&stream));
  tl_stream = stream;
  RAJA::resources::Cuda res = RAJA::resources::Cuda::CudaFromStream(stream);
  RAJA::b200::MultiResource mres(rank, world, ptrs, group->bytes(), res);

  // ---- global data, the same on every rank (each rank uploads only its shard)
  std::vector<int> gi(N);
  std::vector<double> gd(N);
  std::mt19937 gen(12345);
  for (long i = 0; i < N; ++i) { gi[i] = (int)(gen() % 201) - 100; gd[i] = 0.5 + (gen() % 100000) / 100000.0; }
  if (N > 10) { gi[N / 3] = -1000; gi[(2 * N) / 3] = -1000; gi[N / 2] = 1000; gi[N - 1] = 1000; }  // duplicate extrema in different shards

  auto seg       = mres.shard(0L, N);
  const long lo  = *seg.begin(), n = (long)seg.size();
  int* di        = to_dev(gi.data() + lo, (size_t)n);
  double* dd     = to_dev(gd.data() + lo, (size_t)n);
  RAJA::b200::ShardedSpan<int, long> a(di, lo);
  RAJA::b200::ShardedSpan<double, long> x(dd, lo);
  // Ranks are threads of ONE process: cudaMalloc / cudaFree of memory that is mapped into the peers may have to wait for the peer
  // devices, and a peer device may be running an exchange kernel that waits for THIS rank -- a rank that allocates or frees between
  // two collectives can deadlock the group (seen at world 8: the 10 s trap of the waiting kernel).  So: every buffer is allocated
  // before the first collective and freed after the last one, with a host rendezvous on both sides, and the stream-ordered pools
  // (the library's scratch) are warmed by rank-local calls of the same sizes.  (One process per GPU has none of this.)
  double* out; CK(cudaMalloc(&out, (size_t)n * 8 + 64));
  int* so; CK(cudaMalloc(&so, (size_t)n * 4 + 64));
  double* dob; CK(cudaMalloc(&dob, (size_t)n * 8 + 64));
  unsigned* dk; CK(cudaMalloc(&dk, (size_t)n * 4 + 64));
  unsigned* dv; CK(cudaMalloc(&dv, (size_t)n * 4 + 64));
  double* df; CK(cudaMalloc(&df, (size_t)n * 8 + 64));
  if (n > 0) {  // warm-up: local scan / sorts / reduction of the shard's size on this rank's stream
    using LEXEC = RAJA::b200_exec<256>;
    CK(cudaMemsetAsync(dk, 1, (size_t)n * 4, stream)); CK(cudaMemsetAsync(dv, 2, (size_t)n * 4, stream)); CK(cudaMemsetAsync(df, 3, (size_t)n * 8, stream));
    RAJA::inclusive_scan<LEXEC>(res, RAJA::make_span(di, n), RAJA::make_span(so, n), RAJA::operators::plus<int> {});
    RAJA::inclusive_scan<LEXEC>(res, RAJA::make_span(dd, n), RAJA::make_span(dob, n), RAJA::operators::plus<double> {});
    RAJA::sort_pairs<LEXEC>(res, RAJA::make_span(dk, n), RAJA::make_span(dv, n));
    RAJA::sort<LEXEC>(res, RAJA::make_span(df, n));
    long wsum = 0;
    RAJA::forall<LEXEC>(res, RAJA::TypedRangeSegment<long>(0, n), RAJA::expt::Reduce<RAJA::operators::plus>(&wsum),
                        [=] RAJA_DEVICE(long i, RAJA::expt::ValOp<long, RAJA::operators::plus>& s_) { s_ += di[i]; });
  }
  res.wait();
  group->rendezvous();

  // ---- forall + reducer parameters across ranks
  {
    long sum = 7; int mn = 5000, mx = -5000; double dot = 1.0;
    using VL = RAJA::expt::ValLoc<int, long>;
    VL minloc(5000, -1), maxloc(-5000, -1);
    RAJA::forall<MEXEC>(mres.resource(), seg, RAJA::expt::Reduce<RAJA::operators::plus>(&sum), RAJA::expt::Reduce<RAJA::operators::minimum>(&mn),
                        RAJA::expt::Reduce<RAJA::operators::maximum>(&mx), RAJA::expt::Reduce<RAJA::operators::plus>(&dot),
                        RAJA::expt::Reduce<RAJA::operators::minimum>(&minloc), RAJA::expt::Reduce<RAJA::operators::maximum>(&maxloc),
                        [=] RAJA_DEVICE(long i, RAJA::expt::ValOp<long, RAJA::operators::plus>& s, RAJA::expt::ValOp<int, RAJA::operators::minimum>& lo_,
                                        RAJA::expt::ValOp<int, RAJA::operators::maximum>& hi_, RAJA::expt::ValOp<double, RAJA::operators::plus>& d,
                                        RAJA::expt::ValLocOp<int, long, RAJA::operators::minimum>& ml, RAJA::expt::ValLocOp<int, long, RAJA::operators::maximum>& xl) {
                          s += a[i]; lo_.min(a[i]); hi_.max(a[i]); d += x[i] * x[i]; ml.minloc(a[i], i); xl.maxloc(a[i], i);
                        });
    long rsum = 7; int rmn = 5000, rmx = -5000; double rdot = 1.0, rabs = 0; long rminloc = -1, rmaxloc = -1; int rminv = 5000, rmaxv = -5000;
    for (long i = 0; i < N; ++i) {
      rsum += gi[i]; rmn = std::min(rmn, gi[i]); rmx = std::max(rmx, gi[i]); rdot += gd[i] * gd[i]; rabs += gd[i] * gd[i];
      if (gi[i] < rminv) { rminv = gi[i]; rminloc = i; }
      if (gi[i] > rmaxv) { rmaxv = gi[i]; rmaxloc = i; }
    }
    EXPECT(sum == rsum); EXPECT(mn == rmn); EXPECT(mx == rmx);
    EXPECT(std::fabs(dot - rdot) <= N * 2.3e-16 * rabs);
    EXPECT(minloc.getVal() == rminv); EXPECT(minloc.getLoc() == rminloc);  // first extremum, also across shard boundaries
    EXPECT(maxloc.getVal() == rmaxv); EXPECT(maxloc.getLoc() == rmaxloc);
  }
  // ---- a plain sharded forall (no exchange): triad on the shard with global indices
  {
    RAJA::b200::ShardedSpan<double, long> o(out, lo);
    RAJA::forall<MEXEC>(mres.resource(), seg, [=] RAJA_DEVICE(long i) { o[i] = x[i] + 3.0 * (double)i; });
    auto h = to_host(out, (size_t)n);
    bool ok = true;
    for (long i = 0; i < n; ++i) ok = ok && (h[i] == gd[lo + i] + 3.0 * (double)(lo + i));
    EXPECT(ok);
  }
  // ---- scans with the carry exchanged between ranks
  {
    std::vector<int> ref(N);
    RAJA::inclusive_scan<RAJA::seq_exec>(RAJA::make_span(gi.data(), N), RAJA::make_span(ref.data(), N), RAJA::operators::plus<int> {});
    // RAJA's front-end returns early on an empty range (pattern/scan.hpp:72): a rank with an empty shard uses the collective entry point
    if (n > 0) RAJA::inclusive_scan<MEXEC>(mres.resource(), RAJA::make_span(di, n), RAJA::make_span(so, n), RAJA::operators::plus<int> {});
    else RAJA::b200::multi_inclusive_scan(mres, di, so, n, RAJA::operators::plus<int> {});
    EXPECT(to_host(so, (size_t)n) == std::vector<int>(ref.begin() + lo, ref.begin() + lo + n));
    RAJA::exclusive_scan<RAJA::seq_exec>(RAJA::make_span(gi.data(), N), RAJA::make_span(ref.data(), N), RAJA::operators::minimum<int> {}, 77);
    if (n > 0) RAJA::exclusive_scan<MEXEC>(mres.resource(), RAJA::make_span(di, n), RAJA::make_span(so, n), RAJA::operators::minimum<int> {}, 77);
    else RAJA::b200::multi_exclusive_scan(mres, di, so, n, RAJA::operators::minimum<int> {}, 77);
    EXPECT(to_host(so, (size_t)n) == std::vector<int>(ref.begin() + lo, ref.begin() + lo + n));
    copy_sync(dob, dd, (size_t)n * 8, cudaMemcpyDeviceToDevice);
    std::vector<double> rd(gd);
    RAJA::inclusive_scan_inplace<RAJA::seq_exec>(RAJA::make_span(rd.data(), N), RAJA::operators::plus<double> {});
    if (n > 0) RAJA::inclusive_scan_inplace<MEXEC>(mres.resource(), RAJA::make_span(dob, n), RAJA::operators::plus<double> {});
    else RAJA::b200::multi_inclusive_scan(mres, dob, dob, n, RAJA::operators::plus<double> {});
    auto h = to_host(dob, (size_t)n);
    bool ok = true;
    for (long i = 0; i < n; ++i) ok = ok && (std::fabs(h[i] - rd[lo + i]) <= N * 2.3e-16 * rd[N - 1]);
    EXPECT(ok);
  }
  // ---- sample sort across ranks: u32 keys (every third key from a 256-value alphabet: long runs of duplicates that straddle
  //      ranks) + the global index as value; fp64 keys only, descending
  {
    std::vector<unsigned> gk(N), gv(N);
    for (long i = 0; i < N; ++i) { gk[i] = (i % 3 == 0) ? (gen() & 0xffu) : gen(); gv[i] = (unsigned)i; }
    copy_sync(dk, gk.data() + lo, (size_t)n * 4, cudaMemcpyHostToDevice);
    copy_sync(dv, gv.data() + lo, (size_t)n * 4, cudaMemcpyHostToDevice);
    auto piece = RAJA::b200::multi_sort_pairs(mres, dk, dv, (long long)n);
    mres.resource().wait();
    col->k32[rank] = to_host(piece.keys, (size_t)piece.n);
    col->v32[rank] = to_host(piece.vals, (size_t)piece.n);
    group->rendezvous();
    if (rank == 0) {
      std::vector<unsigned> ak, av;
      for (int r = 0; r < world; ++r) { ak.insert(ak.end(), col->k32[r].begin(), col->k32[r].end()); av.insert(av.end(), col->v32[r].begin(), col->v32[r].end()); }
      RAJA::stable_sort_pairs<RAJA::seq_exec>(RAJA::make_span(gk.data(), N), RAJA::make_span(gv.data(), N));
      EXPECT(ak == gk); EXPECT(av == gv);
    }
    group->rendezvous();
    std::vector<double> gf(N);
    for (long i = 0; i < N; ++i) gf[i] = ((gen() % 2000001) - 1000000.0) * (i % 7 == 0 ? 0.0 : 1.0) * (i % 11 == 0 ? -1.0 : 1.0);
    copy_sync(df, gf.data() + lo, (size_t)n * 8, cudaMemcpyHostToDevice);
    auto pf = RAJA::b200::multi_sort(mres, df, (long long)n, /*descending=*/true);
    mres.resource().wait();
    col->f64[rank] = to_host(pf.keys, (size_t)pf.n);
    group->rendezvous();
    if (rank == 0) {
      std::vector<double> af;
      for (int r = 0; r < world; ++r) af.insert(af.end(), col->f64[r].begin(), col->f64[r].end());
      RAJA::stable_sort<RAJA::seq_exec>(RAJA::make_span(gf.data(), N), RAJA::operators::greater<double> {});
      EXPECT(af.size() == gf.size());
      bool ok = af.size() == gf.size();
      for (size_t i = 0; ok && i < af.size(); ++i) ok = (af[i] == gf[i]) && (std::signbit(af[i]) == std::signbit(gf[i]));  // +-0.0 keep their input order
      EXPECT(ok);
    }
    group->rendezvous();
  }
  mres.barrier();
  mres.resource().wait();
  group->rendezvous();  // nobody frees while a peer may still be inside a collective
  cudaFree(di); cudaFree(dd); cudaFree(out); cudaFree(so); cudaFree(dob); cudaFree(dk); cudaFree(dv); cudaFree(df);
  group->rendezvous();
}

int main(int argc, char** argv)
{
  // Ranks are THREADS of one process here, and the exchange kernels wait for their peers on the device.  With lazy module
  // loading the first launch of a kernel may have to synchronise the context it is loaded into -- behind a kernel that is
  // itself waiting for a peer whose host thread is stuck the same way.  Load everything up front.  (One process per GPU,
  // the deployment mode of bench.py, has one context per process and does not need this.)
  setenv("CUDA_MODULE_LOADING", "EAGER", 1);
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  int world = ndev < 8 ? ndev : 8;
  if (const char* w = getenv("RB200_WORLD")) world = atoi(w);
  // RB200_OVERSUBSCRIBE=1: more ranks than devices (rank r runs on device r % ndev; ranks that share a device exchange through plain
  // device memory) -- the world-8 logic (seven splitters, empty shards) on a smaller box
  if (world < 1 || (world > ndev && !getenv("RB200_OVERSUBSCRIBE"))) { printf("test_multi: world %d needs that many devices (%d visible)\n", world, ndev); return 2; }
  std::vector<long> sizes = {1000003L, 5L, 40000L};  // 5 elements over up to 8 ranks: empty shards
  if (const char* one = getenv("RB200_TEST_N")) sizes = {atol(one)};
  for (long N : sizes) {
    RAJA::b200::LocalPeerGroup group(world, RAJA::b200::MultiResource::workspace_bytes((size_t)(N / world + 1) * 3 + 65536, 8, 4));
    Collect col;
    col.k32.resize(world); col.v32.resize(world); col.f64.resize(world);
    std::vector<std::thread> th;
    for (int r = 0; r < world; ++r)
      th.emplace_back([&, r] {
        try {
          rank_main(r, world, &group, &col, N);
        } catch (const std::exception& e) {  // say what failed before the process goes down (the other ranks are waiting for this one)
          printf("test_multi: rank %d of %d, N = %ld: exception: %s\n", r, world, N, e.what());
          fflush(stdout);
          _exit(3);
        }
      });
    for (auto& t : th) t.join();
    printf("test_multi: N = %ld world = %d: %s\n", N, world, g_fail ? "FAILED" : "ok");
  }
  if (g_fail) { printf("test_multi: %d failure(s)\n", g_fail); return 1; }
  printf("test_multi: OK (world = %d)\n", world);
  return 0;
}
