// Permutohedral lattice on the device: construction (embedding, hash table with 128-bit CAS, vertex
// numbering, blur neighbours, vertex-major CSR of the splat matrix) and the three filter stages
// splat / blur / slice.  Reference: third-party/densecrf/src/permutohedral.cpp:54-131 (hash table),
// :140-321 (init, SSE build), :476-589 (seqCompute / sseCompute).
//
// Design notes (B200):
//  * Embedding arithmetic mirrors the SSE code operation by operation (separate float multiplies and adds,
//    round-half-even), so every point lands in the same simplex with the same barycentric weights as in the
//    reference.  The vertex NUMBERING differs (it is not observable); the partition of points is identical.
//  * Keys (d x int16) are packed into one 128-bit word and claimed with a single ATOMG.CAS.128.  A slot goes
//    EMPTY -> key exactly once, so plain cached loads are safe for the fast "already there" path.
//  * Splat is a gather, not a scatter: the (point, corner) pairs are counting-sorted by vertex once per
//    lattice, rows are cut into segments of SPLAT_SEG nonzeros, and each iteration every (segment, float4
//    channel group) work item sums its w * Q rows and issues one vector RED.ADD.F32x4 per item.  This removes
//    the atomic hot spot of the few-vertices regime (a 640x480 frame at the node's kernel widths has ~10^3
//    vertices for 2*10^6 adds per channel) while staying load-balanced in the many-vertices regime.
//  * No kernel needs the vertex count on the host: V lives in device memory, launches are sized by capacity.
#include "lattice.cuh"
#include "meanfield.cuh"
#include "scan.cuh"

namespace rss {

// ---------------------------------------------------------------------------------------------- keys
__device__ __forceinline__ bool key_eq(const Key128& a, const Key128& b) { return a.lo == b.lo && a.hi == b.hi; }
__device__ __forceinline__ Key128 key_empty() { return Key128{~0ull, ~0ull}; }
__device__ __forceinline__ uint32_t key_hash(const Key128& k) {
    unsigned long long h = k.lo * 0x9E3779B97F4A7C15ull;
    h ^= (k.hi + 0x7F4A7C159E3779B9ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    return (uint32_t)h;
}
template <int D>
__device__ __forceinline__ Key128 key_pack(const int* k) {  // int16 wrap like the reference's short keys
    Key128 r{0ull, 0ull};
#pragma unroll
    for (int i = 0; i < D; i++) {
        const unsigned long long v = (unsigned long long)(k[i] & 0xFFFF);
        if (i < 4) r.lo |= v << (16 * i);
        else r.hi |= v << (16 * (i - 4));
    }
    return r;
}
template <int D>
__device__ __forceinline__ void key_unpack(const Key128& r, int* k) {
#pragma unroll
    for (int i = 0; i < D; i++) {
        const unsigned long long v = i < 4 ? (r.lo >> (16 * i)) : (r.hi >> (16 * (i - 4)));
        k[i] = (int)(short)(v & 0xFFFF);
    }
}
__device__ __forceinline__ Key128 load_key(const Key128* p) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
    return Key128{v.x, v.y};
}
// returns the slot of `key`, inserting it if absent; -1 when the table is full.
// A slot goes EMPTY -> key exactly once.  A plain (cached, possibly stale or torn) 128-bit load may only be trusted
// when it shows OUR key (then the slot was fully written); every other outcome is confirmed with the CAS itself,
// whose return value is the authoritative content of the slot.
// The probe sequence is bounded (HASH_MAX_PROBES): a table that is too small for the lattice must fail FAST (the host
// retries with a larger one), not degenerate into full-table scans by millions of threads.
constexpr uint32_t HASH_MAX_PROBES = 256;
__device__ __forceinline__ int hash_insert(Key128* table, uint32_t mask, const Key128& key) {
    uint32_t h = key_hash(key) & mask;
    const uint32_t limit = min(mask, HASH_MAX_PROBES);
    for (uint32_t probes = 0; probes <= limit; probes++) {
        if (key_eq(load_key(table + h), key)) return (int)h;
        const Key128 old = atomicCAS(table + h, key_empty(), key);
        if (key_eq(old, key_empty())) return (int)h;
        if (key_eq(old, key)) return (int)h;
        h = (h + 1) & mask;
    }
    return -1;
}
__device__ __forceinline__ int hash_find(const Key128* table, uint32_t mask, const Key128& key) {
    uint32_t h = key_hash(key) & mask;
    const uint32_t limit = min(mask, HASH_MAX_PROBES);  // same bound as hash_insert: a key is never farther from home
    for (uint32_t probes = 0; probes <= limit; probes++) {
        const Key128 cur = load_key(table + h);
        if (key_eq(cur, key)) return (int)h;
        if (key_eq(cur, key_empty())) return -1;
        h = (h + 1) & mask;
    }
    return -1;
}

// ---------------------------------------------------------------------------------------------- embedding
struct ScaleFactors {
    float s[LAT_MAX_D];
};

// Permutohedral::init, SSE build (permutohedral.cpp:192-277), one thread per point.
template <int D>
__global__ void __launch_bounds__(256) lattice_embed_kernel(const float* __restrict__ feat, int N, int Next, ScaleFactors sf,
                                                            Key128* __restrict__ table, uint32_t mask,
                                                            int* __restrict__ offsets, float* __restrict__ bary_out,
                                                            uint32_t* __restrict__ first_ref,
                                                            uint32_t* __restrict__ counts) {
    __shared__ int s_off[256 * (D + 1)];
    __shared__ float s_bary[256 * (D + 1)];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    // no early exit: the lanes of a warp agree on duplicate vertices below (full-mask match / shuffle)
    // (an L2 read, not a system-scope volatile one: a stale 0 only means this thread still works on a build that is void)
    const bool live = i < Next && !__ldcg(counts + 1);  // overflow: the build is void anyway
    // i == N (only when N % 4 != 0): the reference's SSE loop pads its last block of four points with ZERO features and
    // still inserts their vertices (permutohedral.cpp:192-198,268-275).  Those vertices never receive a splat, but they
    // exist for the blur and pass values on between their neighbours - so they must exist here too.
    const bool padding = i >= N;
    const float invdplus1 = __fdiv_rn(1.0f, (float)(D + 1)), dplus1 = (float)(D + 1);
    float elevated[D + 1], rem0[D + 1], rank[D + 1], bary[D + 2];
    // elevate (:203-209)
    float sm = 0.f;
#pragma unroll
    for (int j = D; j > 0; j--) {
        const float cf = __fmul_rn((padding || !live) ? 0.0f : feat[(size_t)i * D + (j - 1)], sf.s[j - 1]);
        elevated[j] = __fsub_rn(sm, __fmul_rn((float)j, cf));
        sm = __fadd_rn(sm, cf);
    }
    elevated[0] = sm;
    // closest 0-coloured simplex (:212-222), cvtps_epi32 = round-half-even
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k <= D; k++) {
        const float v = rintf(__fmul_rn(invdplus1, elevated[k]));
        rem0[k] = __fmul_rn(v, dplus1);
        sum = __fadd_rn(sum, v);
    }
    // rank (:225-235)
#pragma unroll
    for (int k = 0; k <= D; k++) rank[k] = 0.f;
#pragma unroll
    for (int a = 0; a < D; a++) {
        const float da = __fsub_rn(elevated[a], rem0[a]);
#pragma unroll
        for (int b = a + 1; b <= D; b++) {
            const float db = __fsub_rn(elevated[b], rem0[b]);
            const float c = da < db ? 1.f : 0.f;
            rank[a] = __fadd_rn(rank[a], c);
            rank[b] = __fadd_rn(rank[b], __fsub_rn(1.f, c));
        }
    }
    // bring the point back onto the plane (:238-244)
#pragma unroll
    for (int k = 0; k <= D; k++) {
        rank[k] = __fadd_rn(rank[k], sum);
        const float add = rank[k] < 0.f ? dplus1 : 0.f, sub = rank[k] >= dplus1 ? dplus1 : 0.f;
        const float adj = __fsub_rn(add, sub);
        rank[k] = __fadd_rn(rank[k], adj);
        rem0[k] = __fadd_rn(rem0[k], adj);
    }
    // barycentric coordinates (:247-265)
#pragma unroll
    for (int k = 0; k < D + 2; k++) bary[k] = 0.f;
#pragma unroll
    for (int k = 0; k <= D; k++) {
        const float v = __fmul_rn(__fsub_rn(elevated[k], rem0[k]), invdplus1);
        const int q = D - (int)rank[k];
#pragma unroll
        for (int z = 0; z < D + 2; z++) {  // register-resident scatter
            if (z == q) bary[z] = __fadd_rn(bary[z], v);
            if (z == q + 1) bary[z] = __fsub_rn(bary[z], v);
        }
    }
    bary[0] = __fadd_rn(bary[0], __fadd_rn(1.f, bary[D + 1]));
    // vertices of the simplex (:270-277): key = rem0 + canonical[remainder][rank]
    int irank[D + 1], irem[D + 1];
#pragma unroll
    for (int k = 0; k <= D; k++) {
        irank[k] = (int)rank[k];
        irem[k] = __float2int_rz(rem0[k]);
    }
#pragma unroll
    for (int rem = 0; rem <= D; rem++) {
        int key[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            // canonical[rem][r] = rem if r <= D - rem else rem - (D+1)   (:171-177)
            const int canon = irank[k] <= D - rem ? rem : rem - (D + 1);
            key[k] = irem[k] + canon;
        }
        // Neighbouring pixels land on the same vertices: the lanes of the warp that hold the SAME key elect the lowest one,
        // which alone probes / claims the hash slot and records the first (point, corner) pair (it has the smallest point
        // index of the group); the others get the slot by shuffle.  Dead lanes carry a key nobody else can have.
        Key128 kk = key_pack<D>(key);
        if (!live) { kk.lo = 0xFFFF000000000000ull | (unsigned long long)(threadIdx.x & 31); kk.hi = ~0ull; }
        unsigned grp = __match_any_sync(0xffffffffu, kk.lo);
        if (D > 4) grp &= __match_any_sync(0xffffffffu, kk.hi);
        const int leader = __ffs(grp) - 1;
        int slot = -1;
        if (live && (int)(threadIdx.x & 31) == leader) {
            slot = hash_insert(table, mask, kk);
            if (slot < 0) counts[1] = 1u;
            else atomicMin(first_ref + slot, (uint32_t)i * (D + 1) + rem);  // first (point, corner) pair that touches the vertex
        }
        slot = __shfl_sync(0xffffffffu, slot, leader);
        // staged in shared memory and written out below: a thread's D+1 values are 4 (D+1) bytes apart from its neighbour's,
        // so storing them one corner at a time scatters every 4-byte store into its own sector (9e6 L2 sectors for D = 5)
        s_off[threadIdx.x * (D + 1) + rem] = slot;
        s_bary[threadIdx.x * (D + 1) + rem] = bary[rem];
    }
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * blockDim.x * (D + 1);
    const size_t end = (size_t)Next * (D + 1);
#pragma unroll
    for (int k = 0; k <= D; k++) {  // (also after an overflow: the host rebuilds everything then)
        const int q = k * 256 + threadIdx.x;
        if (base + q < end) {
            offsets[base + q] = s_off[q];
            bary_out[base + q] = s_bary[q];
        }
    }
}

// Vertex numbering in order of first appearance in the point sequence - the numbering the reference's serial hash
// table produces (permutohedral.cpp:118-120).  Besides making offsets comparable with the reference one to one, it
// is what makes the filter cache friendly: consecutive points (neighbouring pixels) share vertices, so vertices that
// are close in the lattice get close ids, and the value rows a warp gathers sit in a few cache lines.
// Computed WITHOUT a pass over the (point, corner) pairs: every occupied hash slot marks the index of its first
// pair (first_ref, an atomicMin of the embedding kernel) in a bitmap over pair indices (one thread per SLOT: 10^5, not 10^6 pairs), one CTA turns the bitmap's word
// popcounts into prefix sums, and every slot reads its rank = vertices whose first pair comes earlier.
__global__ void __launch_bounds__(256) first_bitmap_kernel(const uint32_t* __restrict__ first_ref, uint32_t hcap,
                                                           const uint32_t* __restrict__ counts, uint32_t* __restrict__ bitmap) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= hcap || counts[1]) return;
    const uint32_t r = first_ref[s];
    if (r != 0xFFFFFFFFu) atomicOr(bitmap + (r >> 5), 1u << (r & 31));
}
// prefix[w] = number of set bits in words [0, w): two launches, every load coalesced and independent.  A CTA owns a chunk
// of 4096 words (1024 threads x one uint4); pass 1 writes the chunks' popcounts, pass 2 adds up the chunks before its own
// (a few hundred words even for a 15 M-point map) and scans its chunk.  nwords is a multiple of 4 (the host rounds up).
constexpr uint32_t BMP_CHUNK = 4096;
__device__ __forceinline__ uint32_t block_sum_1024(uint32_t v, uint32_t* wsum) {  // result valid in every thread
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = __reduce_add_sync(0xffffffffu, v);
    if (lane == 0) wsum[wid] = v;
    __syncthreads();
    uint32_t t = wsum[lane];
    t = __reduce_add_sync(0xffffffffu, t);
    __syncthreads();
    return t;
}
__global__ void __launch_bounds__(1024) bitmap_chunk_sums_kernel(const uint32_t* __restrict__ bitmap, uint32_t nwords,
                                                                 uint32_t* __restrict__ sums) {
    __shared__ uint32_t wsum[32];
    const uint32_t w = blockIdx.x * BMP_CHUNK + threadIdx.x * 4;
    uint32_t s = 0;
    if (w < nwords) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(bitmap + w));
        s = __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
    s = block_sum_1024(s, wsum);
    if (threadIdx.x == 0) sums[blockIdx.x] = s;
}
__global__ void __launch_bounds__(1024) bitmap_prefix_kernel(const uint32_t* __restrict__ bitmap, uint32_t nwords,
                                                             const uint32_t* __restrict__ sums, uint32_t* __restrict__ prefix,
                                                             uint32_t* __restrict__ counts, uint32_t vcap) {
    __shared__ uint32_t wsum[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t w = blockIdx.x * BMP_CHUNK + threadIdx.x * 4;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (w < nwords) v = __ldg(reinterpret_cast<const uint4*>(bitmap + w));
    uint32_t before = 0;
    for (uint32_t c = threadIdx.x; c < blockIdx.x; c += blockDim.x) before += __ldg(sums + c);
    before = block_sum_1024(before, wsum);
    const uint32_t p0 = __popc(v.x), p1 = __popc(v.y), p2 = __popc(v.z), p3 = __popc(v.w), s = p0 + p1 + p2 + p3;
    uint32_t inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const uint32_t x = wsum[lane];
        uint32_t ix = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, ix, o);
            if (lane >= o) ix += t;
        }
        wsum[lane] = ix - x;
        if (lane == 31 && blockIdx.x == gridDim.x - 1) {  // total = number of vertices
            const uint32_t V = before + ix;
            counts[0] = V;
            if (V > vcap) counts[1] = 1u;  // load factor above 1/2: the host retries with a bigger table
        }
    }
    __syncthreads();
    const uint32_t run = before + wsum[wid] + inc - s;
    if (w < nwords) *reinterpret_cast<uint4*>(prefix + w) = make_uint4(run, run + p0, run + p0 + p1, run + p0 + p1 + p2);
}
__global__ void __launch_bounds__(256) assign_ids_bitmap_kernel(const uint32_t* __restrict__ first_ref, uint32_t hcap,
                                                                const uint32_t* __restrict__ bitmap,
                                                                const uint32_t* __restrict__ prefix,
                                                                const Key128* __restrict__ table, uint32_t vcap,
                                                                uint32_t* __restrict__ slot_id, Key128* __restrict__ vkeys,
                                                                const uint32_t* __restrict__ counts) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= hcap || counts[1]) return;
    const uint32_t r = first_ref[s];
    if (r == 0xFFFFFFFFu) return;
    const uint32_t id = prefix[r >> 5] + __popc(bitmap[r >> 5] & ((1u << (r & 31)) - 1u));
    slot_id[s] = id;
    if (id < vcap) vkeys[id] = load_key(table + s);
}
__global__ void __launch_bounds__(256) remap_offsets_kernel(int* __restrict__ offsets, size_t n,
                                                            const uint32_t* __restrict__ slot_id,
                                                            uint32_t* __restrict__ deg, const uint32_t* __restrict__ counts,
                                                            uint32_t vcap) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || counts[1]) return;
    const int slot = offsets[k];
    const uint32_t id = slot_id[slot];
    offsets[k] = (int)id;
    if (deg && id < vcap) atomicAdd(deg + id, 1u);
}

// counts[5] += number of (point, corner) pairs whose vertex differs from the same corner of the previous point: the
// number of atomics the fused mean-field splat (meanfield.cu) would issue per channel group with unbounded chunks
__global__ void __launch_bounds__(256) run_count_kernel(const int* __restrict__ offsets, size_t n, int d1,
                                                        uint32_t* __restrict__ counts) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool differs = !counts[1] && k < n && (k < (size_t)d1 || offsets[k] != offsets[k - d1]);
    const unsigned m = __ballot_sync(0xffffffffu, differs);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(counts + 5, (uint32_t)__popc(m));
}
void launch_run_count(rss_ctx* c, cudaStream_t st, Lattice& L) {
    const size_t nnz = (size_t)L.N * (L.d + 1);
    RSS_LAUNCH(c, run_count_kernel, rss_div_up((long long)nnz, 256), 256, 0, st, L.offsets.as<int>(), nnz, L.d + 1,
               L.counts.as<uint32_t>());
}

// blur neighbours (:303-318): along axis j, n1 = key - 1 with coordinate j set to key[j] + d, n2 the opposite
template <int D>
__global__ void __launch_bounds__(256) neighbors_kernel(const Key128* __restrict__ vkeys, const Key128* __restrict__ table,
                                                        uint32_t mask, const uint32_t* __restrict__ slot_id,
                                                        const uint32_t* __restrict__ counts, uint32_t vcap,
                                                        int2* __restrict__ nbr) {
    const uint32_t V = counts[0];
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (counts[1] || gid >= V * (D + 1)) return;
    const uint32_t j = gid / V, v = gid - j * V;
    int key[D], n1[D], n2[D];
    key_unpack<D>(vkeys[v], key);
#pragma unroll
    for (int k = 0; k < D; k++) { n1[k] = key[k] - 1; n2[k] = key[k] + 1; }
#pragma unroll
    for (int k = 0; k < D; k++)
        if ((uint32_t)k == j) { n1[k] = key[k] + D; n2[k] = key[k] - D; }
    const int s1 = hash_find(table, mask, key_pack<D>(n1)), s2 = hash_find(table, mask, key_pack<D>(n2));
    nbr[(size_t)j * vcap + v] = make_int2(s1 < 0 ? (int)vcap : (int)slot_id[s1], s2 < 0 ? (int)vcap : (int)slot_id[s2]);
}

// vertex-major CSR of the splat matrix + segment list
__global__ void __launch_bounds__(256) seg_count_kernel(const uint32_t* __restrict__ deg, const uint32_t* __restrict__ counts,
                                                        uint32_t* __restrict__ nseg, uint32_t vcap) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > vcap) return;
    const uint32_t V = counts[1] ? 0u : counts[0];
    nseg[v] = v < V ? (deg[v] + SPLAT_SEG - 1) / SPLAT_SEG : 0u;
}
__global__ void __launch_bounds__(256) csr_fill_kernel(const int* __restrict__ offsets, const float* __restrict__ bary,
                                                       size_t n, int d1, const uint32_t* __restrict__ row_start,
                                                       uint32_t* __restrict__ cursor, const uint32_t* __restrict__ counts,
                                                       int* __restrict__ csr_pt, float* __restrict__ csr_w) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || counts[1]) return;
    const int v = offsets[k];
    const uint32_t pos = row_start[v] + atomicAdd(cursor + v, 1u);
    csr_pt[pos] = (int)(k / d1);
    csr_w[pos] = bary[k];
}
__global__ void __launch_bounds__(256) seg_fill_kernel(const uint32_t* __restrict__ row_start,
                                                       const uint32_t* __restrict__ seg_off,
                                                       const uint32_t* __restrict__ counts, uint32_t maxseg,
                                                       int* __restrict__ seg_v, uint32_t* __restrict__ seg_begin,
                                                       uint32_t* __restrict__ seg_end) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (counts[1] || v >= counts[0]) return;
    const uint32_t b = row_start[v], e = row_start[v + 1];
    uint32_t s = seg_off[v];
    for (uint32_t p = b; p < e && s < maxseg; p += SPLAT_SEG, s++) {
        seg_v[s] = (int)v;
        seg_begin[s] = p;
        seg_end[s] = min(e, p + SPLAT_SEG);
    }
}

// ---------------------------------------------------------------------------------------------- filter stages
// splat (:545-553): values[v] += w * in[p] for every (p, corner) pair of vertex v.  `in` rows are in_stride floats
// (a multiple of 4, 16-byte aligned), optionally scaled by norm[p] first (DenseKernel::filter, pairwise.cpp:65-66).
// One warp per segment of <= 32 nonzeros: the lanes first load the segment's (point, weight, norm) triples with
// three coalesced loads, then the warp walks the nonzeros four at a time - 8 lanes per nonzero, lane g of a group
// reading the float4 channel group g of that point's row, i.e. one contiguous 16*G-byte read per nonzero.
// Partial sums are combined with two xor-shuffles and leave as one RED.ADD.F32x4 per channel group.
__global__ void __launch_bounds__(512) splat_kernel(const int* __restrict__ seg_v, const uint32_t* __restrict__ seg_begin,
                                                    const uint32_t* __restrict__ seg_end, const uint32_t* __restrict__ counts,
                                                    const int* __restrict__ csr_pt, const float* __restrict__ csr_w,
                                                    const float* __restrict__ in, int in_stride,
                                                    const float* __restrict__ norm, int G, int Mp,
                                                    float* __restrict__ values) {
    // Persistent CTAs, each sweeping one CONTIGUOUS range of segments: consecutive segments belong to consecutive
    // vertices (numbered by first appearance = spatially coherent), which reference the same points, so the Q rows
    // fetched for one vertex are still in this SM's L1 when the neighbouring vertices need them.
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    if (counts[1]) return;
    const uint32_t nseg = counts[2];
    const uint32_t per = (nseg + gridDim.x - 1) / gridDim.x;
    const uint32_t s_begin = blockIdx.x * per, s_end = min(nseg, s_begin + per);
    const int sub = lane >> 3, g = lane & 7;
    for (uint32_t warp = s_begin + wib; warp < s_end; warp += wpb) {
    const uint32_t b = seg_begin[warp], e = seg_end[warp];
    const int n = (int)(e - b);
    int p = 0;
    float w = 0.f, nv = 1.f;
    if (lane < n) {
        p = __ldg(csr_pt + b + lane);
        w = __ldg(csr_w + b + lane);
        if (norm) nv = __ldg(norm + p);
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < SPLAT_SEG / 4; it++) {
        const int idx = it * 4 + sub;
        const int pp = __shfl_sync(0xffffffffu, p, idx);
        const float ww = __shfl_sync(0xffffffffu, w, idx);
        const float nn = __shfl_sync(0xffffffffu, nv, idx);
        if (idx < n && g < G) {
            float4 x = __ldg(reinterpret_cast<const float4*>(in + (size_t)pp * in_stride) + g);
            if (norm) {
                x.x = __fmul_rn(x.x, nn); x.y = __fmul_rn(x.y, nn); x.z = __fmul_rn(x.z, nn); x.w = __fmul_rn(x.w, nn);
            }
            acc.x = __fadd_rn(acc.x, __fmul_rn(ww, x.x));
            acc.y = __fadd_rn(acc.y, __fmul_rn(ww, x.y));
            acc.z = __fadd_rn(acc.z, __fmul_rn(ww, x.z));
            acc.w = __fadd_rn(acc.w, __fmul_rn(ww, x.w));
        }
    }
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
        acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (lane < G) {
        float* dst = values + (size_t)seg_v[warp] * Mp + 4 * lane;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(acc.x), "f"(acc.y), "f"(acc.z),
                     "f"(acc.w)
                     : "memory");
    }
    }
}

// blur along one axis (:555-569): new[v] = old[v] + 0.5 * (old[n1] + old[n2]); a missing neighbour is the zero row
// one launch per axis: the large-lattice path (value tables beyond L2 reach of one cluster)
__global__ void __launch_bounds__(256) blur_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                                   const int2* __restrict__ nbr, const uint32_t* __restrict__ counts,
                                                   int G, int vcap) {
    const uint32_t V = counts[0];
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (counts[1] || gid >= (long long)V * G) return;
    const uint32_t v = (uint32_t)(gid / G);
    const int g = (int)(gid - (long long)v * G);
    const int2 nb = __ldg(nbr + v);
    // a missing neighbour is the zero row (index vcap): skip the load, the row is hot enough to serialise on one L2 slice
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    dst[(size_t)v * G + g] = blur_item(src[(size_t)v * G + g], nb.x != vcap ? src[(size_t)nb.x * G + g] : z,
                                       nb.y != vcap ? src[(size_t)nb.y * G + g] : z);
}
__global__ void __launch_bounds__(256) zero_rows_kernel(float4* __restrict__ p, const uint32_t* __restrict__ counts, int G) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < (long long)counts[0] * G) p[gid] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// All d+1 axes in ONE launch for lattices whose value table is L2 resident (the single-frame regime: 10^3..10^5
// vertices, where a per-axis launch is pure launch latency).  A persistent grid of co-resident CTAs (cooperative
// launch, one CTA per SM) walks the table; between axes the CTAs meet at a grid barrier built on one L2 counter
// (monotonic target, so it never needs resetting).  Value reads use ld.global.cg: the rows were written by other SMs
// one axis earlier and must come from L2, not from a stale L1 line.  After the last axis the table that does NOT hold
// the result is zeroed: it is the next iteration's splat target, so an iteration needs no memset launch.
__global__ void __launch_bounds__(512) blur_coop_kernel(float4* __restrict__ a, float4* __restrict__ b,
                                                        const int2* __restrict__ nbr, const uint32_t* __restrict__ counts,
                                                        int G, int d1, uint32_t vcap, unsigned int* barrier,
                                                        unsigned int barrier_base, int reverse) {
    const uint32_t V = counts[1] ? 0u : counts[0];
    const uint32_t items = V * (uint32_t)G;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    float4* src = a;
    float4* dst = b;
    for (int j = 0; j < d1; j++) {
        // reverse: axes d .. 0, the transposed filter (Permutohedral::compute(out, in, true), permutohedral.cpp:555)
        const int2* nb_j = nbr + (size_t)(reverse ? d1 - 1 - j : j) * vcap;
        for (uint32_t it = tid; it < items; it += nthr) {
            const uint32_t v = it / (uint32_t)G, g = it - v * (uint32_t)G;
            const int2 nb = __ldg(nb_j + v);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 o = __ldcg(src + it), x = nb.x != (int)vcap ? __ldcg(src + (size_t)nb.x * G + g) : z,
                         y = nb.y != (int)vcap ? __ldcg(src + (size_t)nb.y * G + g) : z;
            __stcg(dst + it, blur_item(o, x, y));
        }
        grid_barrier(barrier, barrier_base + (unsigned int)(j + 1) * gridDim.x);
        float4* t = src; src = dst; dst = t;
    }
    // src holds the result; dst becomes the next splat target
    for (uint32_t it = tid; it < items; it += nthr) __stcg(dst + it, make_float4(0.f, 0.f, 0.f, 0.f));
}

// plain slice (:571-584) for rss_crf_filter and the normalisation pass; seq selects the scalar path's
// (w * value) * alpha association (seqCompute :518-521) instead of (w * alpha) * value.
__global__ void __launch_bounds__(256) slice_kernel(const int* __restrict__ offsets, const float* __restrict__ bary, int N,
                                                    int d1, float alpha, const float* __restrict__ values, int M, int Mp,
                                                    int seq, const uint32_t* __restrict__ counts,
                                                    float* __restrict__ out, int out_stride) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (counts[1] || gid >= (long long)N * M) return;
    const int i = (int)(gid / M), c = (int)(gid - (long long)i * M);
    float acc = 0.f;
    for (int j = 0; j < d1; j++) {
        const int v = offsets[(size_t)i * d1 + j];
        const float w = bary[(size_t)i * d1 + j];
        const float val = values[(size_t)v * Mp + c];
        const float p = seq ? __fmul_rn(__fmul_rn(w, val), alpha) : __fmul_rn(__fmul_rn(w, alpha), val);
        acc = __fadd_rn(acc, p);
    }
    out[(size_t)i * out_stride + c] = acc;
}

// DenseKernel::initLattice (pairwise.cpp:44-61): norm from the filter response to all-ones
__global__ void __launch_bounds__(256) norm_kernel(float* __restrict__ norm, int N, int norm_type) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float v = norm[i];
    if (norm_type == RSS_NORMALIZE_SYMMETRIC) norm[i] = (float)(1.0 / sqrt((double)v + 1e-20));
    else norm[i] = (float)(1.0 / ((double)v + 1e-20));
}
__global__ void __launch_bounds__(256) fill_f32_kernel(float* __restrict__ p, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void __launch_bounds__(256) finalize_counts_kernel(uint32_t* counts, const uint32_t* seg_total,
                                                              uint32_t vcap, uint32_t maxseg) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (counts[0] > vcap) counts[1] = 1u;
        uint32_t s = *seg_total;
        if (s > maxseg) { counts[1] = 1u; s = 0; }
        counts[2] = counts[1] ? 0u : s;
    }
}

// ---------------------------------------------------------------------------------------------- host side
static ScaleFactors make_scale_factors(int d) {
    ScaleFactors sf;
    const float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (d + 1));                     // permutohedral.cpp:180
    for (int i = 0; i < LAT_MAX_D; i++)
        sf.s[i] = i < d ? (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * inv_std_dev) : 0.f;  // :183
    return sf;
}

template <int D>
static void launch_embed(rss_ctx* c, cudaStream_t st, const float* feat, int N, Lattice& L) {
    const int Next = N + (N % 4 ? 1 : 0);
    RSS_LAUNCH(c, lattice_embed_kernel<D>, rss_div_up(Next, 256), 256, 0, st, feat, N, Next, make_scale_factors(D),
               L.table.as<Key128>(), L.hcap - 1, L.offsets.as<int>(), L.bary.as<float>(), L.first_ref.as<uint32_t>(),
               L.counts.as<uint32_t>());
}
template <int D>
static void launch_neighbors(rss_ctx* c, cudaStream_t st, Lattice& L) {
    RSS_LAUNCH(c, neighbors_kernel<D>, rss_div_up((long long)L.vcap * (D + 1), 256), 256, 0, st, L.vkeys.as<Key128>(),
               L.table.as<Key128>(), L.hcap - 1, L.slot_id.as<uint32_t>(), L.counts.as<uint32_t>(), L.vcap,
               L.nbr.as<int2>());
}

float lattice_alpha(int d) { return 1.0f / (1 + powf(2, -(float)d)); }  // :571

// Allocates for capacity hcap and enqueues the whole construction on `st`.  feat: device [N][d].  Mp = padded
// channel count of the value tables.  No host synchronisation; overflow is reported in counts[1].
rss_status lattice_build(rss_ctx* ctx, cudaStream_t st, Lattice& L, const float* feat, int N, int d, uint32_t hcap, int Mp,
                         bool want_csr) {
    if (d < 1 || d > LAT_MAX_D) return ctx->fail(RSS_ERR_INVALID, "pairwise feature dimension must be in [1, 7]");
    L.d = d; L.N = N; L.hcap = hcap; L.vcap = hcap / 2;
    const int d1 = d + 1;
    const size_t nnz = (size_t)N * d1;                      // pairs of the real points
    const size_t nnz_ext = nnz + (N % 4 ? (size_t)d1 : 0);  // + the reference's zero-feature padding point (see the embed kernel)
    L.maxseg = (uint32_t)(nnz / SPLAT_SEG + L.vcap + 1);
    RSS_CU(ctx, L.table.reserve((size_t)hcap * sizeof(Key128)));
    RSS_CU(ctx, L.slot_id.reserve((size_t)hcap * 4));
    RSS_CU(ctx, L.first_ref.reserve((size_t)hcap * 4));
    RSS_CU(ctx, L.rank.reserve((nnz_ext + 64) * 4));  // bitmap + prefix (nnz / 32 words each, rounded) + chunk sums
    RSS_CU(ctx, L.vkeys.reserve((size_t)L.vcap * sizeof(Key128)));
    RSS_CU(ctx, L.offsets.reserve(nnz_ext * 4));
    RSS_CU(ctx, L.bary.reserve(nnz_ext * 4));
    RSS_CU(ctx, L.nbr.reserve((size_t)d1 * L.vcap * sizeof(int2)));
    RSS_CU(ctx, L.norm.reserve((size_t)N * 4));
    RSS_CU(ctx, L.counts.reserve(64));
    RSS_CU(ctx, L.deg.reserve((size_t)(L.vcap + 2) * 4));
    RSS_CU(ctx, L.cursor.reserve((size_t)(L.vcap + 1) * 4));
    RSS_CU(ctx, L.nseg.reserve((size_t)(L.vcap + 2) * 4));
    if (want_csr) {
        RSS_CU(ctx, L.csr_pt.reserve(nnz * 4));
        RSS_CU(ctx, L.csr_w.reserve(nnz * 4));
        RSS_CU(ctx, L.seg_v.reserve((size_t)L.maxseg * 4));
        RSS_CU(ctx, L.seg_begin.reserve((size_t)L.maxseg * 4));
        RSS_CU(ctx, L.seg_end.reserve((size_t)L.maxseg * 4));
    }
    L.have_csr = want_csr;
    L.tile_TP = 0;
    RSS_CU(ctx, L.val_a.reserve((size_t)(L.vcap + 1) * Mp * 4));
    RSS_CU(ctx, L.val_b.reserve((size_t)(L.vcap + 1) * Mp * 4));
    RSS_CU(ctx, L.val_c.reserve((size_t)(L.vcap + 1) * Mp * 4));
    RSS_CU(ctx, L.scan_tmp.reserve((scan_tmp_elems(nnz_ext > hcap ? nnz_ext : hcap) + 8) * 4));
    uint32_t* counts = L.counts.as<uint32_t>();
    RSS_CU(ctx, cudaMemsetAsync(L.table.ptr, 0xFF, (size_t)hcap * sizeof(Key128), st));
    RSS_CU(ctx, cudaMemsetAsync(L.first_ref.ptr, 0xFF, (size_t)hcap * 4, st));
    RSS_CU(ctx, cudaMemsetAsync(counts, 0, 64, st));
    if (want_csr) {
        RSS_CU(ctx, cudaMemsetAsync(L.deg.ptr, 0, (size_t)(L.vcap + 2) * 4, st));
        RSS_CU(ctx, cudaMemsetAsync(L.cursor.ptr, 0, (size_t)(L.vcap + 1) * 4, st));
    }
    RSS_CU(ctx, cudaMemsetAsync(L.val_a.ptr, 0, (size_t)(L.vcap + 1) * Mp * 4, st));
    RSS_CU(ctx, cudaMemsetAsync(L.val_b.ptr, 0, (size_t)(L.vcap + 1) * Mp * 4, st));
    RSS_CU(ctx, cudaMemsetAsync(L.val_c.ptr, 0, (size_t)(L.vcap + 1) * Mp * 4, st));
    L.splat_target = 0;
    L.barrier_base = 0;
    switch (d) {
        case 1: launch_embed<1>(ctx, st, feat, N, L); break;
        case 2: launch_embed<2>(ctx, st, feat, N, L); break;
        case 3: launch_embed<3>(ctx, st, feat, N, L); break;
        case 4: launch_embed<4>(ctx, st, feat, N, L); break;
        case 5: launch_embed<5>(ctx, st, feat, N, L); break;
        case 6: launch_embed<6>(ctx, st, feat, N, L); break;
        default: launch_embed<7>(ctx, st, feat, N, L); break;
    }
    // vertex numbering by first appearance (bitmap of first pairs -> popcount prefix -> rank per slot); counts[0] = V
    {
        const uint32_t nwords = (uint32_t)(((nnz_ext + 31) / 32 + 3) & ~(size_t)3);  // whole uint4s
        const uint32_t nchunks = (nwords + BMP_CHUNK - 1) / BMP_CHUNK;
        uint32_t* bitmap = L.rank.as<uint32_t>();
        uint32_t* prefix = bitmap + nwords;
        uint32_t* sums = prefix + nwords;
        RSS_CU(ctx, cudaMemsetAsync(bitmap, 0, (size_t)nwords * 4, st));
        RSS_LAUNCH(ctx, first_bitmap_kernel, rss_div_up(hcap, 256), 256, 0, st, L.first_ref.as<uint32_t>(), hcap, counts, bitmap);
        RSS_LAUNCH(ctx, bitmap_chunk_sums_kernel, nchunks, 1024, 0, st, bitmap, nwords, sums);
        RSS_LAUNCH(ctx, bitmap_prefix_kernel, nchunks, 1024, 0, st, bitmap, nwords, sums, prefix, counts, L.vcap);
        RSS_LAUNCH(ctx, assign_ids_bitmap_kernel, rss_div_up(hcap, 256), 256, 0, st, L.first_ref.as<uint32_t>(), hcap, bitmap, prefix,
                   L.table.as<Key128>(), L.vcap, L.slot_id.as<uint32_t>(), L.vkeys.as<Key128>(), counts);
    }
    RSS_LAUNCH(ctx, remap_offsets_kernel, rss_div_up((long long)nnz, 256), 256, 0, st, L.offsets.as<int>(), nnz,  // real pairs only
               L.slot_id.as<uint32_t>(), want_csr ? L.deg.as<uint32_t>() : (uint32_t*)nullptr, counts, L.vcap);
    switch (d) {
        case 1: launch_neighbors<1>(ctx, st, L); break;
        case 2: launch_neighbors<2>(ctx, st, L); break;
        case 3: launch_neighbors<3>(ctx, st, L); break;
        case 4: launch_neighbors<4>(ctx, st, L); break;
        case 5: launch_neighbors<5>(ctx, st, L); break;
        case 6: launch_neighbors<6>(ctx, st, L); break;
        default: launch_neighbors<7>(ctx, st, L); break;
    }
    if (!want_csr) {
        RSS_CU(ctx, cudaGetLastError());
        return RSS_OK;
    }
    // CSR: segments per row, row_start = scan(deg), seg_off = scan(nseg)
    RSS_LAUNCH(ctx, seg_count_kernel, rss_div_up((long long)L.vcap + 1, 256), 256, 0, st, L.deg.as<uint32_t>(), counts,
               L.nseg.as<uint32_t>(), L.vcap);
    exclusive_scan_u32(L.deg.as<uint32_t>(), L.deg.as<uint32_t>(), (size_t)L.vcap + 1, L.scan_tmp.as<uint32_t>(), nullptr,
                       st, &ctx->launches);
    exclusive_scan_u32(L.nseg.as<uint32_t>(), L.nseg.as<uint32_t>(), (size_t)L.vcap + 1, L.scan_tmp.as<uint32_t>(),
                       counts + 4, st, &ctx->launches);
    RSS_LAUNCH(ctx, finalize_counts_kernel, 1, 32, 0, st, counts, counts + 4, L.vcap, L.maxseg);
    RSS_LAUNCH(ctx, csr_fill_kernel, rss_div_up((long long)nnz, 256), 256, 0, st, L.offsets.as<int>(), L.bary.as<float>(),
               nnz, d1, L.deg.as<uint32_t>(), L.cursor.as<uint32_t>(), counts, L.csr_pt.as<int>(), L.csr_w.as<float>());
    RSS_LAUNCH(ctx, seg_fill_kernel, rss_div_up(L.vcap, 256), 256, 0, st, L.deg.as<uint32_t>(), L.nseg.as<uint32_t>(),
               counts, L.maxseg, L.seg_v.as<int>(), L.seg_begin.as<uint32_t>(), L.seg_end.as<uint32_t>());
    RSS_CU(ctx, cudaGetLastError());
    return RSS_OK;
}

// blur_coop_kernel on tables a (input) / b with G float4 items per vertex: result in a when d+1 is even, else in b; the
// other table comes out all zero
static cudaError_t launch_blur_coop(rss_ctx* ctx, cudaStream_t st, Lattice& L, float4* pa, float4* pb, int G, int reverse = 0) {
    const int2* pn = L.nbr.as<int2>();
    const uint32_t* pc = L.counts.as<uint32_t>();
    int Garg = G, d1arg = L.d + 1;
    uint32_t vc = L.vcap;
    unsigned int* bar = L.counts.as<unsigned int>() + 8;  // counts[8]: the barrier word, zeroed at build time
    const int grid = ctx->sm_count;
    unsigned int base = L.barrier_base;
    L.barrier_base += (unsigned int)d1arg * (unsigned int)grid;
    void* args[] = {&pa, &pb, &pn, &pc, &Garg, &d1arg, &vc, &bar, &base, &reverse};
    cudaEvent_t ea = nullptr, eb = nullptr;
    if (ctx->profile) { ea = ctx->prof_event(); eb = ctx->prof_event(); cudaEventRecord(ea, st); }
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)blur_coop_kernel, dim3(grid), dim3(512), args, 0, st);
    ctx->launches++;
    if (ctx->profile) { cudaEventRecord(eb, st); ctx->prof_pending.push_back(rss_ctx::Pending{"blur_coop_kernel", ea, eb}); }
    return e;
}

// splat -> (d+1) blurs; returns the table holding the blurred values (NULL when the launch failed).  in: [N][in_stride] (in_stride % 4 == 0).
// Invariant: L.splat_target is all zero on entry; on exit the other table is (or becomes) the next target.
float* lattice_splat_blur(rss_ctx* ctx, cudaStream_t st, Lattice& L, const float* in, int in_stride, const float* norm,
                          int Mp, bool reverse) {
    const int G = Mp / 4, d1 = L.d + 1;
    float* a = L.splat_target ? L.val_b.as<float>() : L.val_a.as<float>();
    float* b = L.splat_target ? L.val_a.as<float>() : L.val_b.as<float>();
    RSS_LAUNCH(ctx, splat_kernel, ctx->sm_count * 4, 512, 0, st, L.seg_v.as<int>(),
               L.seg_begin.as<uint32_t>(), L.seg_end.as<uint32_t>(), L.counts.as<uint32_t>(), L.csr_pt.as<int>(),
               L.csr_w.as<float>(), in, in_stride, norm, G, Mp, a);
    const bool small = (size_t)L.vcap * G <= (size_t)BLUR_COOP_MAX_ITEMS;
    if (small) {
        if (launch_blur_coop(ctx, st, L, reinterpret_cast<float4*>(a), reinterpret_cast<float4*>(b), G, reverse ? 1 : 0) != cudaSuccess) return nullptr;
    } else {
        float* s = a;
        float* d = b;
        for (int j = 0; j < d1; j++) {
            RSS_LAUNCH(ctx, blur_kernel, rss_div_up((long long)L.vcap * G, 256), 256, 0, st,
                       reinterpret_cast<const float4*>(s), reinterpret_cast<float4*>(d),
                       L.nbr.as<int2>() + (size_t)(reverse ? d1 - 1 - j : j) * L.vcap, L.counts.as<uint32_t>(), G, (int)L.vcap);
            float* t = s; s = d; d = t;
        }
        RSS_LAUNCH(ctx, zero_rows_kernel, rss_div_up((long long)L.vcap * G, 256), 256, 0, st, reinterpret_cast<float4*>(d),
                   L.counts.as<uint32_t>(), G);
    }
    // after d1 swaps the result is in `a` when d1 is even, else in `b`; the other one was zeroed
    float* result = (d1 % 2 == 0) ? a : b;
    float* next_target = (d1 % 2 == 0) ? b : a;
    L.splat_target = next_target == L.val_b.as<float>() ? 1 : 0;
    return result;
}

void lattice_slice(rss_ctx* ctx, cudaStream_t st, Lattice& L, const float* values, int M, int Mp, int seq, float* out,
                   int out_stride) {
    RSS_LAUNCH(ctx, slice_kernel, rss_div_up((long long)L.N * M, 256), 256, 0, st, L.offsets.as<int>(), L.bary.as<float>(),
               L.N, L.d + 1, lattice_alpha(L.d), values, M, Mp, seq, L.counts.as<uint32_t>(), out, out_stride);
}

// splat of the all-ones vector: values[v][0] = sum of the barycentric weights that reference v (no Q gather needed)
__global__ void __launch_bounds__(256) splat_ones_kernel(const int* __restrict__ seg_v, const uint32_t* __restrict__ seg_begin,
                                                         const uint32_t* __restrict__ seg_end, const uint32_t* __restrict__ counts,
                                                         const float* __restrict__ csr_w, float* __restrict__ values) {
    const uint32_t warp = (uint32_t)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (counts[1] || warp >= counts[2]) return;
    const uint32_t b = seg_begin[warp], e = seg_end[warp];
    float w = b + lane < e ? __ldg(csr_w + b + lane) : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (lane == 0) atomicAdd(values + (size_t)seg_v[warp] * 4, w);
}

// norm_ = 1/sqrt(K 1 + 1e-20) through the scalar path (pairwise.cpp:44,54-57; permutohedral.cpp:600-601).
// The value tables (allocated for the CRF's Mp >= 4) are used with a row stride of 4 floats, channel 0 live.
rss_status lattice_normalization(rss_ctx* ctx, cudaStream_t st, Lattice& L) {
    const int N = L.N, d1 = L.d + 1;
    float* norm = L.norm.as<float>();
    // NO_NORMALIZATION: DenseKernel::filter applies norm_ neither before nor after the lattice (pairwise.cpp:65,78); the
    // mean norm it computes (:46-52) only enters the parameter gradients, so nothing is needed here
    if (L.norm_type == RSS_NO_NORMALIZATION) return RSS_OK;
    float* a = L.splat_target ? L.val_b.as<float>() : L.val_a.as<float>();
    float* b = L.splat_target ? L.val_a.as<float>() : L.val_b.as<float>();
    if (L.have_csr)
        RSS_LAUNCH(ctx, splat_ones_kernel, rss_div_up((long long)L.maxseg * 32, 256), 256, 0, st, L.seg_v.as<int>(),
                   L.seg_begin.as<uint32_t>(), L.seg_end.as<uint32_t>(), L.counts.as<uint32_t>(), L.csr_w.as<float>(), a);
    else  // no CSR (keyframe path): run-accumulating scatter over the points
        launch_splat_ones_runs(ctx, st, L.offsets.as<int>(), L.bary.as<float>(), N, d1, L.counts.as<uint32_t>(), a);
    // all d+1 axes in one cooperative launch when the table is small (one float4 item per vertex); it leaves the result in
    // `s` and has already cleared the other table
    float* s = a;
    float* d = b;
    const bool small = (size_t)L.vcap <= (size_t)BLUR_COOP_MAX_ITEMS;
    if (small) {
        RSS_CU(ctx, launch_blur_coop(ctx, st, L, reinterpret_cast<float4*>(a), reinterpret_cast<float4*>(b), 1));
        s = (d1 % 2 == 0) ? a : b;
        d = (d1 % 2 == 0) ? b : a;
    } else {
        for (int j = 0; j < d1; j++) {
            RSS_LAUNCH(ctx, blur_kernel, rss_div_up((long long)L.vcap, 256), 256, 0, st, reinterpret_cast<const float4*>(s),
                       reinterpret_cast<float4*>(d), L.nbr.as<int2>() + (size_t)j * L.vcap, L.counts.as<uint32_t>(), 1, (int)L.vcap);
            float* t = s; s = d; d = t;
        }
        RSS_LAUNCH(ctx, zero_rows_kernel, rss_div_up((long long)L.vcap, 256), 256, 0, st, reinterpret_cast<float4*>(d),
                   L.counts.as<uint32_t>(), 1);
    }
    lattice_slice(ctx, st, L, s, 1, 4, 1, norm, 1);
    RSS_LAUNCH(ctx, norm_kernel, rss_div_up(N, 256), 256, 0, st, norm, N, L.norm_type);
    // the result table is dirty in its first V*4 floats: clear it again for the filter proper
    RSS_LAUNCH(ctx, zero_rows_kernel, rss_div_up((long long)L.vcap, 256), 256, 0, st, reinterpret_cast<float4*>(s),
               L.counts.as<uint32_t>(), 1);
    RSS_CU(ctx, cudaGetLastError());
    return RSS_OK;
}

// per-tile data of the fused mean-field kernel (after the normalisation: pre- / post-scale, Potts weight and slice scale
// are folded into the weights)
rss_status lattice_build_tile_csr(rss_ctx* ctx, cudaStream_t st, Lattice& L, int G, int grid_w, int grid_h, const int* perm) {
    TileMap tm = fused_tile_map(L.N, grid_w, grid_h);
    tm.perm = tm.W == 0 ? perm : nullptr;
    const int d1 = L.d + 1;
    const size_t ntiles = (size_t)tm.ntiles, cap = ntiles * tm.TP * d1;
    RSS_CU(ctx, L.tile_pairs.reserve(cap * sizeof(uint2)));
    RSS_CU(ctx, L.tile_ent_meta.reserve(cap * sizeof(int2)));
    RSS_CU(ctx, L.tile_vert.reserve(cap * sizeof(int)));
    RSS_CU(ctx, L.tile_pt_w.reserve(cap * sizeof(float)));
    RSS_CU(ctx, L.tile_pt_slot.reserve(cap * sizeof(uint16_t)));
    RSS_CU(ctx, L.tile_info.reserve(ntiles * sizeof(int2)));
    const bool pre = L.norm_type == RSS_NORMALIZE_SYMMETRIC || L.norm_type == RSS_NORMALIZE_BEFORE;
    const bool post = L.norm_type == RSS_NORMALIZE_SYMMETRIC || L.norm_type == RSS_NORMALIZE_AFTER;
    TileCsrOut out;
    out.pairs = L.tile_pairs.as<uint2>(); out.ent_meta = L.tile_ent_meta.as<int2>(); out.tile_vert = L.tile_vert.as<int>();
    out.tile_info = L.tile_info.as<int2>(); out.pt_w = L.tile_pt_w.as<float>(); out.pt_slot = L.tile_pt_slot.as<uint16_t>();
    // tmp -= -w * (alpha * filtered) * norm  (densecrf.cpp:126, labelcompatibility.cpp:46-48, permutohedral.cpp:571)
    launch_tile_csr_build(ctx, st, L.offsets.as<int>(), L.bary.as<float>(), L.norm.as<float>(), pre, post,
                          L.potts_w * lattice_alpha(L.d), tm, d1, G * 16, L.counts.as<uint32_t>(), out);
    L.tile_TP = tm.TP;
    L.tile_W = tm.W;
    RSS_CU(ctx, cudaGetLastError());
    return RSS_OK;
}

}  // namespace rss
