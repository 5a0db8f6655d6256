// Probe: is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) bit-identical to a chain of FMAs with k increasing?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu
// Prints, over many random tiles, how many of the 64 outputs match each candidate order.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

__global__ void probe(const double *A, const double *B, const double *C, double *D, int tiles)
{
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= tiles) return;
    const double *a = A + (size_t)t * 32, *b = B + (size_t)t * 32, *c = C + (size_t)t * 64;
    const int g = lane >> 2, q = lane & 3;
    double av = a[g * 4 + q];          // A(row g, k q)   row-major 8x4
    double bv = b[q * 8 + g];          // B(k q, col g)   row-major 4x8
    double c0 = c[g * 8 + 2 * q], c1 = c[g * 8 + 2 * q + 1];
    double d0, d1;
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(av), "d"(bv), "d"(c0), "d"(c1));
    D[(size_t)t * 64 + g * 8 + 2 * q] = d0;
    D[(size_t)t * 64 + g * 8 + 2 * q + 1] = d1;
}

static double rnd() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

int main()
{
    const int tiles = 4096;
    double *hA = (double *)malloc(tiles * 32 * 8), *hB = (double *)malloc(tiles * 32 * 8);
    double *hC = (double *)malloc(tiles * 64 * 8), *hD = (double *)malloc(tiles * 64 * 8);
    srand(1);
    for (int i = 0; i < tiles * 32; ++i) { hA[i] = rnd(); hB[i] = rnd(); }
    for (int i = 0; i < tiles * 64; ++i) hC[i] = rnd() * (i % 3 == 0 ? 1e-3 : 1.0);
    double *dA, *dB, *dC, *dD;
    cudaMalloc(&dA, tiles * 32 * 8); cudaMalloc(&dB, tiles * 32 * 8); cudaMalloc(&dC, tiles * 64 * 8); cudaMalloc(&dD, tiles * 64 * 8);
    cudaMemcpy(dA, hA, tiles * 32 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, tiles * 32 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dC, hC, tiles * 64 * 8, cudaMemcpyHostToDevice);
    probe<<<tiles / 4, 128>>>(dA, dB, dC, dD, tiles);
    cudaError_t e = cudaMemcpy(hD, dD, tiles * 64 * 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    long inc = 0, dec = 0, sumfirst = 0, total = 0;
    for (int t = 0; t < tiles; ++t)
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) {
                const double *a = hA + t * 32 + i * 4, *b = hB + t * 32;
                double c = hC[t * 64 + i * 8 + j], d = hD[t * 64 + i * 8 + j];
                double x = c;
                for (int k = 0; k < 4; ++k) x = __builtin_fma(a[k], b[k * 8 + j], x);
                double y = c;
                for (int k = 3; k >= 0; --k) y = __builtin_fma(a[k], b[k * 8 + j], y);
                double s = 0;
                for (int k = 0; k < 4; ++k) s = __builtin_fma(a[k], b[k * 8 + j], s);
                s += c;
                inc += (memcmp(&x, &d, 8) == 0);
                dec += (memcmp(&y, &d, 8) == 0);
                sumfirst += (memcmp(&s, &d, 8) == 0);
                ++total;
            }
    printf("{\"dmma_m8n8k4_total\": %ld, \"match_fma_chain_k_increasing\": %ld, \"match_k_decreasing\": %ld, \"match_dot_then_add\": %ld}\n",
           total, inc, dec, sumfirst);
    return 0;
}
