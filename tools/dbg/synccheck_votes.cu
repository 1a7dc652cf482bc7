// Stand-alone reproduction attempt for the synccheck report on the block-wide phase votes (profiles/README.md, "Block-wide phase votes"):
// the same loop shape as find_intersection_sync - two __syncthreads_count votes per turn, a block-uniform choice between two phases,
// thread-divergent work inside each phase - with the first turn peelable (hold == false on entry).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o synccheck_votes synccheck_votes.cu
//   compute-sanitizer --tool synccheck ./synccheck_votes [threads] [items]
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__device__ __noinline__ uint32_t leaf_work(uint32_t leaf, uint32_t seed)
{
    // divergent trip count, like a primitive test
    uint32_t acc = seed;
    const uint32_t n = 1u + (mix(leaf ^ seed) & 63u);
    for (uint32_t k = 0; k < n; k++) acc = mix(acc + k);
    return acc;
}

__global__ void __launch_bounds__(512, 2) k_votes(unsigned int* cursor, uint32_t n, unsigned long long* out)
{
    __shared__ uint32_t s_base;
    unsigned long long sum = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_base = atomicAdd(cursor, blockDim.x);
        __syncthreads();
        const uint32_t base = s_base;
        if (base >= n) break;
        const uint32_t i = base + threadIdx.x;
        bool alive = i < n;
        // a per-thread "stack" of pending nodes: a counter and a seed
        uint32_t sp = alive ? 1u + (mix(i) & 15u) : 0u, seed = mix(i * 2654435761u + 12345u);
        uint32_t leaf = 0xFFFFFFFFu;
        for (;;) {
            const bool want = alive && leaf == 0xFFFFFFFFu && sp > 0;
            const int want_n = __syncthreads_count(want), hold_n = __syncthreads_count(leaf != 0xFFFFFFFFu);
            if ((want_n | hold_n) == 0) break;
            if (hold_n * 2 > want_n || want_n == 0) {
                const uint32_t cur = leaf;
                leaf = 0xFFFFFFFFu;
                if (cur != 0xFFFFFFFFu) {
                    const uint32_t r = leaf_work(cur, seed);
                    sum += r;
                    if ((r & 127u) == 0u) { alive = false; sp = 0; }     // "blocked": this thread is done
                }
            } else if (want) {
                --sp;
                seed = mix(seed + 1u);
                if (seed & 1u) leaf = seed >> 8;               // a leaf to test
                else if ((seed & 6u) == 0u && sp < 24u) sp += 2;   // children pushed
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, sum);
}

int main(int argc, char** argv)
{
    const int threads = argc > 1 ? atoi(argv[1]) : 512;
    const uint32_t n = argc > 2 ? (uint32_t)atoi(argv[2]) : 5184u;
    unsigned int* cursor; unsigned long long* out;
    cudaMalloc(&cursor, 4); cudaMalloc(&out, 8);
    cudaMemset(cursor, 0, 4); cudaMemset(out, 0, 8);
    k_votes<<<148 * 2, threads>>>(cursor, n, out);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h = 0;
    cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("threads %d items %u: %s, checksum %llu\n", threads, n, cudaGetErrorString(e), h);
    return e != cudaSuccess;
}
