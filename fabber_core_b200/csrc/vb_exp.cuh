/*
 * vb_exp.cuh - double-precision exp() for the forward-model hooks, built for the FP64-pipe-bound time loop.
 *
 * CUDA's exp(double) costs 14 DFMA + a DSETP + a guarded slow path per call; the biexp pass makes six
 * calls per sample and the FP64 pipe is the bound (profiles/). This one is the classic table method:
 *      x = k * ln2/64 + r,  |r| <= ln2/128,  k = 64 m + j
 *      exp(x) = 2^m * (T_j + (T_j * p(r) + Tlo_j)),   p(r) = r + r^2/2 + r^3/6 + r^4/24 + r^5/120
 * with T_j + Tlo_j = 2^(j/64) to ~106 bits (table generated with 60-digit decimal arithmetic, staged in
 * shared memory): 11 FP64 instructions, no branch. Accuracy: max 0.79 ULP over 3e5 random arguments
 * against 60-digit references, 96.7 % correctly rounded (CUDA's own exp is documented at 1 ULP); checked
 * on the device against exp() in tests/test_gpu_exp.py.
 * Valid for finite x < 708 (no overflow); x < -708 returns 0 where exp() returns a denormal or 0 (< 3e-308 of
 * the amplitude it multiplies - a decayed exponential; one integer compare). The caller checks the argument
 * range once per pass, outside the time loop, and falls back to exp() otherwise (vb_models.cuh).
 */
#pragma once
#include "vb_device.cuh"

namespace fab
{
__device__ const unsigned long long EXP_TAB_HI_BITS[64] = {
    0x3ff0000000000000ull, 0x3ff02c9a3e778061ull, 0x3ff059b0d3158574ull, 0x3ff0874518759bc8ull,
    0x3ff0b5586cf9890full, 0x3ff0e3ec32d3d1a2ull, 0x3ff11301d0125b51ull, 0x3ff1429aaea92de0ull,
    0x3ff172b83c7d517bull, 0x3ff1a35beb6fcb75ull, 0x3ff1d4873168b9aaull, 0x3ff2063b88628cd6ull,
    0x3ff2387a6e756238ull, 0x3ff26b4565e27cddull, 0x3ff29e9df51fdee1ull, 0x3ff2d285a6e4030bull,
    0x3ff306fe0a31b715ull, 0x3ff33c08b26416ffull, 0x3ff371a7373aa9cbull, 0x3ff3a7db34e59ff7ull,
    0x3ff3dea64c123422ull, 0x3ff4160a21f72e2aull, 0x3ff44e086061892dull, 0x3ff486a2b5c13cd0ull,
    0x3ff4bfdad5362a27ull, 0x3ff4f9b2769d2ca7ull, 0x3ff5342b569d4f82ull, 0x3ff56f4736b527daull,
    0x3ff5ab07dd485429ull, 0x3ff5e76f15ad2148ull, 0x3ff6247eb03a5585ull, 0x3ff6623882552225ull,
    0x3ff6a09e667f3bcdull, 0x3ff6dfb23c651a2full, 0x3ff71f75e8ec5f74ull, 0x3ff75feb564267c9ull,
    0x3ff7a11473eb0187ull, 0x3ff7e2f336cf4e62ull, 0x3ff82589994cce13ull, 0x3ff868d99b4492edull,
    0x3ff8ace5422aa0dbull, 0x3ff8f1ae99157736ull, 0x3ff93737b0cdc5e5ull, 0x3ff97d829fde4e50ull,
    0x3ff9c49182a3f090ull, 0x3ffa0c667b5de565ull, 0x3ffa5503b23e255dull, 0x3ffa9e6b5579fdbfull,
    0x3ffae89f995ad3adull, 0x3ffb33a2b84f15fbull, 0x3ffb7f76f2fb5e47ull, 0x3ffbcc1e904bc1d2ull,
    0x3ffc199bdd85529cull, 0x3ffc67f12e57d14bull, 0x3ffcb720dcef9069ull, 0x3ffd072d4a07897cull,
    0x3ffd5818dcfba487ull, 0x3ffda9e603db3285ull, 0x3ffdfc97337b9b5full, 0x3ffe502ee78b3ff6ull,
    0x3ffea4afa2a490daull, 0x3ffefa1bee615a27ull, 0x3fff50765b6e4540ull, 0x3fffa7c1819e90d8ull,
};
__device__ const unsigned long long EXP_TAB_LO_BITS[64] = {
    0x0000000000000000ull, 0xbc719083535b085dull, 0x3c8d73e2a475b465ull, 0x3c6186be4bb284ffull,
    0x3c98a62e4adc610bull, 0x3c403a1727c57b53ull, 0xbc96c51039449b3aull, 0xbc932fbf9af1369eull,
    0xbc819041b9d78a76ull, 0x3c8e5b4c7b4968e4ull, 0x3c9e016e00a2643cull, 0x3c8dc775814a8495ull,
    0x3c99b07eb6c70573ull, 0x3c82bd339940e9d9ull, 0x3c8612e8afad1255ull, 0x3c90024754db41d5ull,
    0x3c86f46ad23182e4ull, 0x3c932721843659a6ull, 0xbc963aeabf42eae2ull, 0xbc75e436d661f5e3ull,
    0x3c8ada0911f09ebcull, 0xbc5ef3691c309278ull, 0x3c489b7a04ef80d0ull, 0x3c73c1a3b69062f0ull,
    0x3c7d4397afec42e2ull, 0xbc94b309d25957e3ull, 0xbc807abe1db13cadull, 0x3c99bb2c011d93adull,
    0x3c96324c054647adull, 0x3c9ba6f93080e65eull, 0xbc9383c17e40b497ull, 0xbc9bb60987591c34ull,
    0xbc9bdd3413b26456ull, 0xbc6bbe3a683c88abull, 0xbc816e4786887a99ull, 0xbc90245957316dd3ull,
    0xbc841577ee04992full, 0x3c705d02ba15797eull, 0xbc9d4c1dd41532d8ull, 0xbc9fc6f89bd4f6baull,
    0x3c96e9f156864b27ull, 0x3c85cc13a2e3976cull, 0xbc675fc781b57ebcull, 0xbc9d185b7c1b85d1ull,
    0x3c7c7c46b071f2beull, 0xbc9359495d1cd533ull, 0xbc9d2f6edb8d41e1ull, 0x3c90fac90ef7fd31ull,
    0x3c97a1cd345dcc81ull, 0xbc62805e3084d708ull, 0xbc75584f7e54ac3bull, 0x3c823dd07a2d9e84ull,
    0x3c811065895048ddull, 0x3c92884dff483cadull, 0x3c7503cbd1e949dbull, 0xbc9cbc3743797a9cull,
    0x3c82ed02d75b3707ull, 0x3c9c2300696db532ull, 0xbc91a5cd4f184b5cull, 0x3c839e8980a9cc8full,
    0xbc9e9c23179c2893ull, 0x3c9dc7f486a4b6b0ull, 0x3c99d3e12dd8a18bull, 0x3c874853f3a5931eull,
};

constexpr int EXP_TAB_DOUBLES = 128; /* hi[64] then lo[64] */

/* block-cooperative copy of the table into shared memory */
FAB_DEV void exp_table_stage(double *smem_tab)
{
    for (int i = threadIdx.x; i < 64; i += blockDim.x)
    {
        smem_tab[i] = __longlong_as_double((long long)EXP_TAB_HI_BITS[i]);
        smem_tab[64 + i] = __longlong_as_double((long long)EXP_TAB_LO_BITS[i]);
    }
}

FAB_DEV double exp_fast(double x, const double *tab)
{
    const double INV_L = 92.33248261689366;        /* 64 / ln2 */
    const double L_HI = 0x1.62e42fef00000p-7;       /* ln2/64, 33 significant bits: k * L_HI is exact */
    const double L_LO = 0x1.473de6af278edp-40;
    const double MAGIC = 6755399441055744.0;        /* 1.5 * 2^52: round-to-nearest-integer trick */
    double kd = fma(x, INV_L, MAGIC);
    const int k = __double2loint(kd);
    kd -= MAGIC;
    double r = fma(kd, -L_HI, x);
    r = fma(kd, -L_LO, r);
    const int j = k & 63, m = k >> 6;
    const double T = tab[j], Tlo = tab[64 + j];
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(r, q, 1.0 / 6.0);
    q = fma(r, q, 0.5);
    q = fma(r, q, 1.0);
    const double p = r * q;
    const double v = fma(T, p, Tlo) + T;
    /* scale by 2^m on the integer pipe: v is in [1, 2), m in [-1022, 1021] for |x| < 708 */
    const double e = __hiloint2double(__double2hiint(v) + (m << 20), __double2loint(v));
    /* x < -708 (a rate that has run away: the term has decayed to nothing): 0. Sign bit set and magnitude bits
     * above those of 708.0, tested on the high word */
    return ((unsigned)__double2hiint(x) > 0xc0862000u) ? 0.0 : e;
}

/* every argument -rate * t the pass will form is finite and < 708 (no overflow); large positive rates are fine */
FAB_DEV bool exp_fast_range_ok(double rate, double t_max)
{
    return rate * t_max > -708.0 && rate < 1e300; /* false for NaN */
}

} // namespace fab
