// Parity of RAJA::launch<LaunchPolicy<b200_launch_t<Async>>> with expt::Reduce / ReduceLoc / Name parameters
// against the reference's own seq_launch_t on the same bodies.
// Patterns follow test/functional/launch/reduce-params/tests/test-launch-basic-param-expt-Reduce{Sum,Min,BitAnd}.hpp
// (rand() % 100 data, segments (0,28), (3,642), (0,2057), two reducers + Name in one launch, repeated launches
// accumulating into the same target), plus what the combine must also survive here: 2-D / 3-D teams and
// threads, team sizes that are not a multiple of 32, thousands of teams, a launch with no reducers and
// dynamic shared memory.
#include "RAJA_b200/RAJA.hpp"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

static int g_fail = 0;
#define EXPECT(cond)                                                         \
  do {                                                                       \
    if (!(cond)) { printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
  } while (0)

template <class T> T* to_dev(const std::vector<T>& h) { T* p; cudaMalloc(&p, h.size() * sizeof(T) + 16); cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice); return p; }

using B200_LAUNCH = RAJA::LaunchPolicy<RAJA::b200_launch_t<false>>;
using B200_ASYNC  = RAJA::LaunchPolicy<RAJA::b200_launch_t<true>>;
using SEQ_LAUNCH  = RAJA::LaunchPolicy<RAJA::seq_launch_t>;
using GLOBAL_X    = RAJA::LoopPolicy<RAJA::cuda_global_thread_x>;
using SEQ_LOOP    = RAJA::LoopPolicy<RAJA::seq_exec>;

template <class T> using VSum = RAJA::expt::ValOp<T, RAJA::operators::plus>;
template <class T> using VMin = RAJA::expt::ValOp<T, RAJA::operators::minimum>;
template <class T> using VMax = RAJA::expt::ValOp<T, RAJA::operators::maximum>;

// the reference's ReduceSum / ReduceMin test shape: 1-D teams of 256 threads over a range segment
template <class T> void test_basic(long first, long last, bool exact)
{
  const long data_len = last;
  std::vector<T> h(data_len);
  double mag = 0;
  for (long i = 0; i < data_len; ++i) {
    h[i] = exact ? (T)(rand() % 100) : (T)((rand() % 2000) / 1000.0 - 1.0);
    mag += std::fabs((double)h[i]);
  }
  T* d = to_dev(h);
  const T* hp = h.data();
  RAJA::TypedRangeSegment<long> seg(first, last);
  constexpr int threads = 256;
  const int blocks = (int)((seg.size() - 1) / threads + 1);
  const double tol = exact ? 0.0 : 4.0 * data_len * (double)std::numeric_limits<T>::epsilon() * (mag + 1);

  T sum(0), sum2(2), mn(1000), mx(-1000);
  T rsum(0), rsum2(2), rmn(1000), rmx(-1000);
  RAJA::launch<B200_LAUNCH>(RAJA::LaunchParams(RAJA::Teams(blocks), RAJA::Threads(threads)),
                            RAJA::expt::Reduce<RAJA::operators::plus>(&sum), RAJA::expt::Reduce<RAJA::operators::plus>(&sum2),
                            RAJA::expt::Reduce<RAJA::operators::minimum>(&mn), RAJA::expt::Reduce<RAJA::operators::maximum>(&mx), RAJA::Name("LaunchBasic"),
                            [=] RAJA_HOST_DEVICE(RAJA::LaunchContext ctx, VSum<T>& s, VSum<T>& s2, VMin<T>& lo, VMax<T>& hi) {
                              RAJA::loop<GLOBAL_X>(ctx, seg, [&](long i) {
                                s += d[i];
                                s2 += d[i];
                                lo.min(d[i]);
                                hi.max(d[i]);
                              });
                            });
  RAJA::launch<SEQ_LAUNCH>(RAJA::LaunchParams(RAJA::Teams(blocks), RAJA::Threads(threads)),
                           RAJA::expt::Reduce<RAJA::operators::plus>(&rsum), RAJA::expt::Reduce<RAJA::operators::plus>(&rsum2),
                           RAJA::expt::Reduce<RAJA::operators::minimum>(&rmn), RAJA::expt::Reduce<RAJA::operators::maximum>(&rmx),
                           [=](RAJA::LaunchContext ctx, VSum<T>& s, VSum<T>& s2, VMin<T>& lo, VMax<T>& hi) {
                             RAJA::loop<SEQ_LOOP>(ctx, seg, [&](long i) {
                               s += hp[i];
                               s2 += hp[i];
                               lo.min(hp[i]);
                               hi.max(hp[i]);
                             });
                           });
  EXPECT(std::fabs((double)sum - (double)rsum) <= tol);
  EXPECT(std::fabs((double)sum2 - (double)rsum2) <= tol);
  EXPECT(mn == rmn);
  EXPECT(mx == rmx);

  // repeated launches accumulate into the target (the reference test's nloops block)
  sum = 0;
  rsum = 0;
  for (int j = 0; j < 2; ++j) {
    RAJA::launch<B200_LAUNCH>(RAJA::LaunchParams(RAJA::Teams(blocks), RAJA::Threads(threads)), RAJA::expt::Reduce<RAJA::operators::plus>(&sum),
                              [=] RAJA_HOST_DEVICE(RAJA::LaunchContext ctx, VSum<T>& s) { RAJA::loop<GLOBAL_X>(ctx, seg, [&](long i) { s += d[i]; }); });
    RAJA::launch<SEQ_LAUNCH>(RAJA::LaunchParams(RAJA::Teams(blocks), RAJA::Threads(threads)), RAJA::expt::Reduce<RAJA::operators::plus>(&rsum),
                             [=](RAJA::LaunchContext ctx, VSum<T>& s) { RAJA::loop<SEQ_LOOP>(ctx, seg, [&](long i) { s += hp[i]; }); });
  }
  EXPECT(std::fabs((double)sum - (double)rsum) <= 2 * tol);
  cudaFree(d);
}

// BitAnd / BitOr (test-launch-basic-param-expt-ReduceBitAnd.hpp)
void test_bits(long N)
{
  std::vector<unsigned> h(N);
  for (long i = 0; i < N; ++i) h[i] = 0xffff0000u | (unsigned)(rand() & 0xff0f) | 0x10u;
  unsigned* d = to_dev(h);
  const unsigned* hp = h.data();
  RAJA::TypedRangeSegment<long> seg(0, N);
  const int blocks = (int)((N - 1) / 128 + 1);
  unsigned a = 0xffffffffu, o = 0, ra = 0xffffffffu, ro = 0;
  using VAnd = RAJA::expt::ValOp<unsigned, RAJA::operators::bit_and>;
  using VOr  = RAJA::expt::ValOp<unsigned, RAJA::operators::bit_or>;
  RAJA::launch<B200_LAUNCH>(RAJA::LaunchParams(RAJA::Teams(blocks), RAJA::Threads(128)), RAJA::expt::Reduce<RAJA::operators::bit_and>(&a),
                            RAJA::expt::Reduce<RAJA::operators::bit_or>(&o), [=] RAJA_HOST_DEVICE(RAJA::LaunchContext ctx, VAnd& x, VOr& y) {
                              RAJA::loop<GLOBAL_X>(ctx, seg, [&](long i) {
                                x &= d[i];
                                y |= d[i];
                              });
                            });
  RAJA::launch<SEQ_LAUNCH>(RAJA::LaunchParams(RAJA::Teams(blocks), RAJA::Threads(128)), RAJA::expt::Reduce<RAJA::operators::bit_and>(&ra),
                           RAJA::expt::Reduce<RAJA::operators::bit_or>(&ro), [=](RAJA::LaunchContext ctx, VAnd& x, VOr& y) {
                             RAJA::loop<SEQ_LOOP>(ctx, seg, [&](long i) {
                               x &= hp[i];
                               y |= hp[i];
                             });
                           });
  EXPECT(a == ra);
  EXPECT(o == ro);
  cudaFree(d);
}

// teams x threads of any shape: every (team, thread) adds one known value; MinLoc / MaxLoc with duplicate extrema
void test_shapes(int tx, int ty, int tz, int bx, int by, int bz)
{
  const long nthreads = (long)bx * by * bz, nteams = (long)tx * ty * tz, N = nthreads * nteams;
  std::vector<int> h(N);
  for (long i = 0; i < N; ++i) h[i] = (rand() % 2001) - 1000;
  // duplicate extrema: the first one must win
  if (N >= 8) { h[N / 3] = -5000; h[N / 3 + N / 4] = -5000; h[N / 5] = 7000; h[N / 2 + 1] = 7000; }
  int* d = to_dev(h);
  long long sum = 0, rsum = 0;
  using VL = RAJA::expt::ValLoc<int, long>;
  using VLMin = RAJA::expt::ValLocOp<int, long, RAJA::operators::minimum>;
  using VLMax = RAJA::expt::ValLocOp<int, long, RAJA::operators::maximum>;
  VL mn(100000, -1), mx(-100000, -1);
  RAJA::launch<B200_ASYNC>(RAJA::LaunchParams(RAJA::Teams(tx, ty, tz), RAJA::Threads(bx, by, bz)), RAJA::expt::Reduce<RAJA::operators::plus>(&sum),
                           RAJA::expt::Reduce<RAJA::operators::minimum>(&mn), RAJA::expt::Reduce<RAJA::operators::maximum>(&mx), RAJA::Name("LaunchShapes"),
                           [=] RAJA_DEVICE(RAJA::LaunchContext, RAJA::expt::ValOp<long long, RAJA::operators::plus>& s, VLMin& lo, VLMax& hi) {
                             const long team = blockIdx.x + gridDim.x * (blockIdx.y + (long)gridDim.y * blockIdx.z);
                             const long thr  = threadIdx.x + blockDim.x * (threadIdx.y + (long)blockDim.y * threadIdx.z);
                             const long i    = team * ((long)blockDim.x * blockDim.y * blockDim.z) + thr;
                             s += (long long)d[i];
                             lo.minloc(d[i], i);
                             hi.maxloc(d[i], i);
                           });
  int rmn = 100000, rmx = -100000;
  long rmnl = -1, rmxl = -1;
  for (long i = 0; i < N; ++i) {
    rsum += h[i];
    if (h[i] < rmn) { rmn = h[i]; rmnl = i; }
    if (h[i] > rmx) { rmx = h[i]; rmxl = i; }
  }
  EXPECT(sum == rsum);
  EXPECT(mn.getVal() == rmn);
  EXPECT(mn.getLoc() == rmnl);
  EXPECT(mx.getVal() == rmx);
  EXPECT(mx.getLoc() == rmxl);
  cudaFree(d);
}

// no reducers: teams + thread loops + dynamic shared memory (the plain LaunchExecute path)
void test_plain(int teams, int threads)
{
  const long N = (long)teams * threads;
  std::vector<int> h(N), out(N);
  for (long i = 0; i < N; ++i) h[i] = rand() % 1000;
  int* d = to_dev(h);
  int* o = to_dev(out);
  using TEAM_X   = RAJA::LoopPolicy<RAJA::cuda_block_x_direct>;
  using THREAD_X = RAJA::LoopPolicy<RAJA::cuda_thread_x_loop>;
  // each team reverses its chunk through shared memory
  RAJA::launch<B200_LAUNCH>(RAJA::LaunchParams(RAJA::Teams(teams), RAJA::Threads(threads), (size_t)threads * sizeof(int)),
                            [=] RAJA_HOST_DEVICE(RAJA::LaunchContext ctx) {
                              RAJA::loop<TEAM_X>(ctx, RAJA::TypedRangeSegment<int>(0, teams), [&](int t) {
                                int* sh = ctx.getSharedMemory<int>(threads);
                                RAJA::loop<THREAD_X>(ctx, RAJA::TypedRangeSegment<int>(0, threads), [&](int k) { sh[k] = d[(long)t * threads + k]; });
                                ctx.teamSync();
                                RAJA::loop<THREAD_X>(ctx, RAJA::TypedRangeSegment<int>(0, threads), [&](int k) { o[(long)t * threads + k] = sh[threads - 1 - k]; });
                                ctx.releaseSharedMemory();
                              });
                            });
  cudaMemcpy(out.data(), o, N * sizeof(int), cudaMemcpyDeviceToHost);
  bool ok = true;
  for (int t = 0; t < teams && ok; ++t)
    for (int k = 0; k < threads; ++k)
      if (out[(long)t * threads + k] != h[(long)t * threads + threads - 1 - k]) { ok = false; break; }
  EXPECT(ok);
  // zero teams: no launch, no error
  RAJA::launch<B200_LAUNCH>(RAJA::LaunchParams(RAJA::Teams(0), RAJA::Threads(threads)), [=] RAJA_HOST_DEVICE(RAJA::LaunchContext) {});
  EXPECT(cudaDeviceSynchronize() == cudaSuccess);
  cudaFree(d); cudaFree(o);
}

int main()
{
  srand(4793);
  const long segs[3][2] = {{0, 28}, {3, 642}, {0, 2057}};
  for (auto& s : segs) {
    test_basic<int>(s[0], s[1], true);
    test_basic<long long>(s[0], s[1], true);
    test_basic<float>(s[0], s[1], false);
    test_basic<double>(s[0], s[1], false);
  }
  test_basic<int>(0, 3000000, true);
  test_basic<double>(5, 3000000, false);
  test_bits(28);
  test_bits(100000);
  test_shapes(1, 1, 1, 1, 1, 1);
  test_shapes(3, 1, 1, 7, 1, 1);     // partial warp only
  test_shapes(5, 4, 3, 7, 5, 3);     // 105 threads: three full warps + 9 lanes
  test_shapes(17, 3, 2, 8, 8, 4);    // 256 threads, 3-D
  test_shapes(6000, 1, 1, 33, 1, 1); // many teams, one full warp + 1 lane
  test_shapes(40, 40, 5, 32, 32, 1); // 8000 teams of 1024
  test_plain(37, 96);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
  printf("test_launch_reduce: %s (%d failures)\n", g_fail ? "FAILED" : "OK", g_fail);
  return g_fail ? 1 : 0;
}
