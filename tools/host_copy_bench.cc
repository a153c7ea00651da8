// micro-benchmark: memcpy vs non-temporal copy, threads x 1 MB pieces (how SetVoxelDataArray stages)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>
#include <thread>
#include <vector>
__attribute__((target("avx2"))) static void nt_copy(float *dst, const float *src, size_t n)
{
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) { dst[i] = src[i]; i++; }
    for (; i + 32 <= n; i += 32)
    {
        __m256i a = _mm256_loadu_si256((const __m256i *)(src + i));
        __m256i b = _mm256_loadu_si256((const __m256i *)(src + i + 8));
        __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 16));
        __m256i d = _mm256_loadu_si256((const __m256i *)(src + i + 24));
        _mm256_stream_si256((__m256i *)(dst + i), a);
        _mm256_stream_si256((__m256i *)(dst + i + 8), b);
        _mm256_stream_si256((__m256i *)(dst + i + 16), c);
        _mm256_stream_si256((__m256i *)(dst + i + 24), d);
    }
    for (; i < n; i++) dst[i] = src[i];
    _mm_sfence();
}
int main(int argc, char **argv)
{
    int threads = argc > 1 ? atoi(argv[1]) : 8;
    size_t n = (size_t)128 * 128 * 128 * 64;
    float *src = (float *)aligned_alloc(4096, n * 4), *dst = (float *)aligned_alloc(4096, n * 4);
    memset(src, 1, n * 4); memset(dst, 0, n * 4);
    size_t piece = 1 << 18, pieces = (n + piece - 1) / piece;
    for (int mode = 0; mode < 2; mode++)
        for (int rep = 0; rep < 4; rep++)
        {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; t++)
                pool.emplace_back([&, t]() {
                    for (size_t p = t; p < pieces; p += threads)
                    {
                        size_t b = p * piece, e = std::min(n, b + piece);
                        if (mode) nt_copy(dst + b, src + b, e - b); else memcpy(dst + b, src + b, (e - b) * 4);
                    }
                });
            for (auto &th : pool) th.join();
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            printf("%s threads %d: %.2f ms  %.1f GB/s\n", mode ? "nt    " : "memcpy", threads, ms, n * 4 / ms / 1e6);
        }
    return 0;
}
