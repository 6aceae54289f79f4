// ysm_chains.cu -- batched loop-closure chain finder on the B200.
//
// Replaces, for a whole batch of query scans at once, the reference's Python
//   GraphSlam.find_possible_loop_closure_chains     yag_slam/graph_slam.py:274-304
//   do_breadth_first_traversal + near_scan_visitor  yag_slam/graph.py:71-98, graph_slam.py:32-39
//   RadiusHashSearch.crude_radius_search            yag_slam/helpers.py:395-431
// so that loop-closure match batches (BASELINE cfg 3: thousands of candidate chains per query set)
// are assembled as the CSR base lists ysm_match_batch takes, without Python loops.
//
// One warp per query vertex q:
//   phase A (k_chain_find, first pass only)  the "near linked" set = vertices reachable from q over
//           graph edges through vertices closer than loop_search_dist (frontier BFS, lanes = frontier
//           entries, visited bitmask + queue in HBM scratch);
//   phase B the reference's candidate walk. Vertex ids are scan numbers, so "sort by num" is the id
//           order: lanes test 32 vertices per step (hash-box test of crude_radius_search, exclusion,
//           exact distance test), the ballots are then consumed in order by the (warp-uniform)
//           chain state machine of graph_slam.py:284-302, including its quirks: squared distance
//           compared with the un-squared loop_search_dist, the last candidate is never examined,
//           an excluded candidate skips the gap test, a trailing partial chain is kept.
// The kernel runs twice: a counting pass sizes the output, the second pass writes members.
//
// Compile with -fmad=false (build.py): dx*dx + dy*dy must round as the reference's
// (dx)**2 + (dy)**2 does.
#include "ysm_internal.h"
#include "../../include/ysm.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <exception>
#include <string>
#include <vector>

namespace {

thread_local std::string t_cerr;

int chain_fail(int code, const std::string& msg) {
  t_cerr = msg;
  return code;
}

#define CCK(x)                                                                                   \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess) {                                                                     \
      rc = chain_fail(YSM_ECUDA, std::string(#x) + ": " + cudaGetErrorString(e_));               \
      goto done;                                                                                 \
    }                                                                                            \
  } while (0)

constexpr int kChainCap = 1024;  // largest loop_search_min_chain_size (the running chain never gets longer)

// Device scratch of the calling thread, grown on demand and kept between calls (a chain search is a
// few hundred microseconds of kernels; per-call cudaMalloc / cudaFree would cost milliseconds).
// Two arenas: inputs + per-query state, and the result lists sized after the counting pass.
struct Arena {
  char* p = nullptr;
  size_t cap = 0, used = 0;
  int device = -1;
  cudaError_t reserve(int dev, size_t bytes) {
    used = 0;
    if (dev == device && bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0; device = dev;
    const size_t want = bytes + bytes / 4 + 4096;
    const cudaError_t e = cudaMalloc((void**)&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  template <typename T> T* take(size_t count) {
    const size_t b = (count * sizeof(T) + 255) & ~(size_t)255;
    T* r = reinterpret_cast<T*>(p + used);
    used += b;
    return r;
  }
};
thread_local Arena t_in, t_out;
inline size_t a256(size_t b) { return (b + 255) & ~(size_t)255; }

struct ChainArgs {
  int n, nq, q0;             // vertices, queries of this chunk, first query of the chunk
  int nw;                    // visited words per query
  int min_chain, fill;
  double dist, res, r2, near_sq;
  const double2* pose;
  const double2* hash;
  const int* adj_ptr;
  const int* adj_idx;
  const int* query;
  unsigned* vis;             // [chunk][nw]
  int* queue;                // [chunk][n]
  int2* counts;              // [nq_total] (chains, members) per query
  const int* chain_base;     // [nq_total] exclusive scans (fill pass)
  const int* member_base;
  int* chain_len;            // [total chains]
  int* members;              // [total members]
};

__global__ void __launch_bounds__(128)
k_chain_find(ChainArgs A) {
  __shared__ int s_tail[4];
  __shared__ int s_chain[4][kChainCap];  // the running (not yet accepted) chain of each warp's query
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ql = blockIdx.x * 4 + warp;  // query within the chunk
  if (ql >= A.nq) return;
  const int qi = A.q0 + ql;
  const int q = A.query[qi];
  const double2 qp = A.pose[q];
  unsigned* vis = A.vis + (size_t)ql * A.nw;
  // ---- phase A: near-linked set (graph.py:71-98 with the visitor of graph_slam.py:32-39) ----------
  if (!A.fill) {
    int* queue = A.queue + (size_t)ql * A.n;
    for (int w = lane; w < A.nw; w += 32) vis[w] = 0u;
    __syncwarp();
    // the start vertex passes its own visitor test iff 0 < distance^2
    int head = 0;
    if (lane == 0) {
      s_tail[warp] = 0;
      if (0.0 < A.near_sq) {
        atomicOr(&vis[q >> 5], 1u << (q & 31));
        queue[0] = q;
        s_tail[warp] = 1;
      }
    }
    __syncwarp();
    while (true) {
      const int tail = s_tail[warp];
      __syncwarp();
      if (head >= tail) break;
      for (int i = head + lane; i < tail; i += 32) {
        const int v = queue[i];
        for (int e = A.adj_ptr[v]; e < A.adj_ptr[v + 1]; e++) {
          const int u = A.adj_idx[e];
          const unsigned bit = 1u << (u & 31);
          if (__ldcg(&vis[u >> 5]) & bit) continue;  // L2: the bits are set with atomics, L1 lines may be stale
          const double2 p = A.pose[u];
          const double dx = qp.x - p.x, dy = qp.y - p.y;
          if (dx * dx + dy * dy < A.near_sq) {
            const unsigned old = atomicOr(&vis[u >> 5], bit);
            if (!(old & bit)) queue[atomicAdd(&s_tail[warp], 1)] = u;
          }
        }
      }
      head = tail;
      __syncwarp();
    }
  }
  // ---- phase B: candidate walk (graph_slam.py:280-302) ----------------------------------------------
  int chain_len = 0, n_ch = 0, n_mem = 0;
  int prev = -1;
  bool prev_excl = false, prev_close = false;
  int* mem_out = A.fill ? A.members + A.member_base[qi] : nullptr;
  int* len_out = A.fill ? A.chain_len + A.chain_base[qi] : nullptr;
  for (int base = 0; base < A.n; base += 32) {
    const int v = base + lane;
    bool cand = false, excl = false, close = false;
    if (v < A.n) {
      const double2 h = A.hash[v];
      const long long kx = (long long)(h.x / A.res), ky = (long long)(h.y / A.res);  // int(): toward zero
      const double bx = (double)kx * A.res, by = (double)ky * A.res;
      const double ex = bx - qp.x, ey = by - qp.y;
      cand = ex * ex + ey * ey < A.r2;
      if (cand) {
        const double2 p = A.pose[v];
        const double dx = qp.x - p.x, dy = qp.y - p.y;
        close = dx * dx + dy * dy <= A.dist;  // squared vs un-squared: reference quirk, kept
        excl = v == q || ((__ldcg(&vis[v >> 5]) >> (v & 31)) & 1u);  // L2 read (set by atomics in phase A)
      }
    }
    unsigned cm = __ballot_sync(0xffffffffu, cand);
    const unsigned em = __ballot_sync(0xffffffffu, excl), clm = __ballot_sync(0xffffffffu, close);
    while (cm) {
      const int b = __ffs(cm) - 1;
      cm &= cm - 1;
      const int c = base + b;
      if (prev >= 0) {  // the pair (v1 = prev, v2 = c)
        if (prev_excl) {
          chain_len = 0;
        } else {
          if (prev_close) {
            // a running chain may still be dropped, so it is kept aside (writing it to the output
            // early would spill past this query's share when the chain is dropped)
            if (lane == 0) s_chain[warp][chain_len] = prev;
            chain_len++;
          }
          if (chain_len >= A.min_chain) {
            __syncwarp();
            if (mem_out)
              for (int i = lane; i < chain_len; i += 32) mem_out[n_mem + i] = s_chain[warp][i];
            __syncwarp();
            if (len_out && lane == 0) len_out[n_ch] = chain_len;
            n_ch++;
            n_mem += chain_len;
            chain_len = 0;
          }
          if (c - prev > 1) chain_len = 0;
        }
      }
      prev = c;
      prev_excl = (em >> b) & 1u;
      prev_close = (clm >> b) & 1u;
    }
  }
  if (chain_len > 0) {
    __syncwarp();
    if (mem_out)
      for (int i = lane; i < chain_len; i += 32) mem_out[n_mem + i] = s_chain[warp][i];
    if (len_out && lane == 0) len_out[n_ch] = chain_len;
    n_ch++;
    n_mem += chain_len;
  }
  if (!A.fill && lane == 0) A.counts[qi] = make_int2(n_ch, n_mem);
}

}  // namespace

struct ysm_chains {
  int device = 0;
  int nq = 0;
  std::vector<int> query_chain_ptr, chain_ptr, members;
  int launches = 0;
  float kernel_ms = 0.f;
};

extern "C" const char* ysm_chains_last_error(void) { return t_cerr.c_str(); }

extern "C" void ysm_chains_destroy(ysm_chains* c) { delete c; }

extern "C" int ysm_chains_get_counts(const ysm_chains* c, int32_t* n_chains, int32_t* n_members, int32_t* launches,
                                     double* kernel_ms) {
  if (!c) return chain_fail(YSM_EINVAL, "ysm_chains_get_counts: null handle");
  if (n_chains) *n_chains = (int32_t)c->chain_ptr.size() - 1;
  if (n_members) *n_members = (int32_t)c->members.size();
  if (launches) *launches = c->launches;
  if (kernel_ms) *kernel_ms = (double)c->kernel_ms;
  return YSM_OK;
}

extern "C" int ysm_chains_copy(const ysm_chains* c, int32_t* query_chain_ptr, int32_t* chain_ptr, int32_t* members) {
  if (!c) return chain_fail(YSM_EINVAL, "ysm_chains_copy: null handle");
  if (query_chain_ptr) std::copy(c->query_chain_ptr.begin(), c->query_chain_ptr.end(), query_chain_ptr);
  if (chain_ptr) std::copy(c->chain_ptr.begin(), c->chain_ptr.end(), chain_ptr);
  if (members) std::copy(c->members.begin(), c->members.end(), members);
  return YSM_OK;
}

static int chains_find_impl(const ysm_chain_query* in, int device, void* stream, ysm_chains** out);

extern "C" int ysm_chains_find(const ysm_chain_query* in, int device, void* stream, ysm_chains** out) {
  ysm_quiesce_device(device);  // (a resident latency kernel would make the allocations / syncs below wait)
  try {  // no exception crosses the C ABI
    return chains_find_impl(in, device, stream, out);
  } catch (const std::exception& e) {
    if (out) *out = nullptr;
    return chain_fail(YSM_ENOMEM, std::string("ysm_chains_find: ") + e.what());
  }
}

static int chains_find_impl(const ysm_chain_query* in, int device, void* stream, ysm_chains** out) {
  if (!in || !out) return chain_fail(YSM_EINVAL, "ysm_chains_find: null argument");
  *out = nullptr;
  const int n = in->n_vertices, nq = in->n_queries;
  if (n <= 0 || nq < 0 || !in->pose_xy || !in->adj_ptr || (nq > 0 && !in->query_vertex))
    return chain_fail(YSM_EINVAL, "ysm_chains_find: empty graph or null arrays");
  if (in->min_chain_size < 1) return chain_fail(YSM_EINVAL, "ysm_chains_find: min_chain_size must be >= 1");
  if (in->min_chain_size > kChainCap) return chain_fail(YSM_EUNSUP, "ysm_chains_find: min_chain_size above 1024");
  if (!(in->loop_search_dist > 0.0)) return chain_fail(YSM_EINVAL, "ysm_chains_find: loop_search_dist must be positive");
  const int ne = in->adj_ptr[n];
  if (ne < 0 || (ne > 0 && !in->adj_idx)) return chain_fail(YSM_EINVAL, "ysm_chains_find: bad adjacency");
  for (int v = 0; v < n; v++)
    if (in->adj_ptr[v + 1] < in->adj_ptr[v]) return chain_fail(YSM_EINVAL, "ysm_chains_find: adj_ptr must be non-decreasing");
  for (int e = 0; e < ne; e++)
    if (in->adj_idx[e] < 0 || in->adj_idx[e] >= n) return chain_fail(YSM_EINVAL, "ysm_chains_find: adj_idx out of range");
  for (int i = 0; i < nq; i++)
    if (in->query_vertex[i] < 0 || in->query_vertex[i] >= n)
      return chain_fail(YSM_EINVAL, "ysm_chains_find: query vertex out of range");
  {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return chain_fail(YSM_ECUDA, std::string("ysm_chains_find: ") + cudaGetErrorString(e));
  }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = YSM_OK;
  ysm_chains* c = new ysm_chains();
  c->device = device;
  c->nq = nq;
  double2 *d_pose = nullptr, *d_hash = nullptr;
  int *d_aptr = nullptr, *d_aidx = nullptr, *d_query = nullptr, *d_queue = nullptr, *d_cbase = nullptr, *d_mbase = nullptr;
  int *d_clen = nullptr, *d_mem = nullptr;
  unsigned* d_vis = nullptr;
  int2* d_counts = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  const int nw = (n + 31) / 32;
  // queries per chunk: the BFS queues take n ints per query (<= 256 MB of scratch)
  const int chunk = (int)std::max<long long>(1, std::min<long long>(std::max(nq, 1), (64ll << 20) / std::max(n, 1)));
  std::vector<int2> counts((size_t)std::max(nq, 1));
  std::vector<int> cbase((size_t)std::max(nq, 1)), mbase((size_t)std::max(nq, 1)), clen;
  long long tot_c = 0, tot_m = 0;
  ChainArgs A;

  c->query_chain_ptr.assign((size_t)nq + 1, 0);
  c->chain_ptr.assign(1, 0);
  if (nq == 0) {
    *out = c;
    return YSM_OK;
  }
  CCK(cudaEventCreate(&ev0));
  CCK(cudaEventCreate(&ev1));
  CCK(t_in.reserve(device, 2 * a256(16 * (size_t)n) + a256(4 * (size_t)(n + 1)) + a256(4 * (size_t)std::max(ne, 1)) +
                               3 * a256(4 * (size_t)nq) + a256(8 * (size_t)nq) + a256(4 * (size_t)nw * (size_t)nq) +
                               a256(4 * (size_t)n * (size_t)chunk)));
  d_pose = t_in.take<double2>((size_t)n);
  d_hash = t_in.take<double2>((size_t)n);
  d_aptr = t_in.take<int>((size_t)n + 1);
  d_aidx = t_in.take<int>((size_t)std::max(ne, 1));
  d_query = t_in.take<int>((size_t)nq);
  d_counts = t_in.take<int2>((size_t)nq);
  d_cbase = t_in.take<int>((size_t)nq);
  d_mbase = t_in.take<int>((size_t)nq);
  d_vis = t_in.take<unsigned>((size_t)nw * (size_t)nq);  // kept across the two passes
  d_queue = t_in.take<int>((size_t)n * (size_t)chunk);
  CCK(cudaMemcpyAsync(d_pose, in->pose_xy, 16 * (size_t)n, cudaMemcpyHostToDevice, st));
  CCK(cudaMemcpyAsync(d_hash, in->hash_xy ? in->hash_xy : in->pose_xy, 16 * (size_t)n, cudaMemcpyHostToDevice, st));
  CCK(cudaMemcpyAsync(d_aptr, in->adj_ptr, 4 * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
  if (ne) CCK(cudaMemcpyAsync(d_aidx, in->adj_idx, 4 * (size_t)ne, cudaMemcpyHostToDevice, st));
  CCK(cudaMemcpyAsync(d_query, in->query_vertex, 4 * (size_t)nq, cudaMemcpyHostToDevice, st));
  A.n = n; A.nw = nw; A.min_chain = in->min_chain_size;
  A.dist = in->loop_search_dist; A.res = in->loop_search_dist;  // RadiusHashSearch(res=loop_search_dist), graph_slam.py:67
  A.r2 = in->crude_r2; A.near_sq = in->near_dist_sq;
  A.pose = d_pose; A.hash = d_hash; A.adj_ptr = d_aptr; A.adj_idx = d_aidx; A.query = d_query;
  A.queue = d_queue; A.counts = d_counts; A.chain_base = d_cbase; A.member_base = d_mbase;
  A.chain_len = nullptr; A.members = nullptr;
  CCK(cudaEventRecord(ev0, st));
  // pass 1: near sets + counts
  A.fill = 0;
  for (int q0 = 0; q0 < nq; q0 += chunk) {
    A.q0 = q0; A.nq = std::min(chunk, nq - q0);
    A.vis = d_vis + (size_t)q0 * nw;
    k_chain_find<<<(A.nq + 3) / 4, 128, 0, st>>>(A);
    c->launches++;
  }
  CCK(cudaGetLastError());
  CCK(cudaMemcpyAsync(counts.data(), d_counts, 8 * (size_t)nq, cudaMemcpyDeviceToHost, st));
  CCK(cudaStreamSynchronize(st));
  for (int i = 0; i < nq; i++) {
    cbase[i] = (int)tot_c; mbase[i] = (int)tot_m;
    tot_c += counts[i].x; tot_m += counts[i].y;
    c->query_chain_ptr[(size_t)i + 1] = (int)tot_c;
  }
  if (tot_c > 0x7fffff00ll || tot_m > 0x7fffff00ll) {
    rc = chain_fail(YSM_EUNSUP, "ysm_chains_find: result exceeds 2^31 entries");
    goto done;
  }
  clen.assign((size_t)std::max<long long>(tot_c, 1), 0);
  c->members.assign((size_t)tot_m, 0);
  if (tot_c > 0) {
    CCK(t_out.reserve(device, a256(4 * (size_t)tot_c) + a256(4 * (size_t)std::max<long long>(tot_m, 1))));
    d_clen = t_out.take<int>((size_t)tot_c);
    d_mem = t_out.take<int>((size_t)std::max<long long>(tot_m, 1));
    CCK(cudaMemcpyAsync(d_cbase, cbase.data(), 4 * (size_t)nq, cudaMemcpyHostToDevice, st));
    CCK(cudaMemcpyAsync(d_mbase, mbase.data(), 4 * (size_t)nq, cudaMemcpyHostToDevice, st));
    // pass 2: write chain lengths and members
    A.fill = 1; A.chain_len = d_clen; A.members = d_mem;
    for (int q0 = 0; q0 < nq; q0 += chunk) {
      A.q0 = q0; A.nq = std::min(chunk, nq - q0);
      A.vis = d_vis + (size_t)q0 * nw;
      k_chain_find<<<(A.nq + 3) / 4, 128, 0, st>>>(A);
      c->launches++;
    }
    CCK(cudaGetLastError());
    CCK(cudaMemcpyAsync(clen.data(), d_clen, 4 * (size_t)tot_c, cudaMemcpyDeviceToHost, st));
    if (tot_m) CCK(cudaMemcpyAsync(c->members.data(), d_mem, 4 * (size_t)tot_m, cudaMemcpyDeviceToHost, st));
  }
  CCK(cudaEventRecord(ev1, st));
  CCK(cudaStreamSynchronize(st));
  CCK(cudaEventElapsedTime(&c->kernel_ms, ev0, ev1));
  c->chain_ptr.assign((size_t)tot_c + 1, 0);
  for (long long i = 0; i < tot_c; i++) c->chain_ptr[(size_t)i + 1] = c->chain_ptr[(size_t)i] + clen[(size_t)i];
done:
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (rc != YSM_OK) {
    delete c;
    return rc;
  }
  *out = c;
  return YSM_OK;
}
