// Exact decimal <-> double decisions for the canonicaliser, no binary floating point:
//   ckf_nearest_double_ex   the double a decimal literal denotes (normal, subnormal, infinity, zero)
//   ckf_is_repr             "is this literal what Python's repr() prints for the double it parses to?"
//   ckf_shortest            the literal repr() would print instead
//
// pydantic-core re-emits a float as the shortest digit string that round-trips (ryu; identical to repr(float)),
// so a float literal in a record is a fixed point of dump(validate(.)) iff, with v = m * 10^k its exact value and
// d = the double nearest to v (ties to even):
//   (B) neither neighbour of v on the 10^(k+1) grid (one digit fewer) rounds to d      -> no shorter spelling exists
//   (C) v is the point of the 10^k grid nearest to d                                   -> it is the one repr picks
// (v rounds to d by construction).  Everything is decided with exact integer arithmetic on small bignums
// (<= 1344 bits: |decimal exponent| <= 345), no floating point, no tables: these literals are rare (computed
// values such as 0.30000000000000004, values at the edges of the range), the common short ones never get here (up to 15
// digits in the normal range the DBL_DIG argument in ck_walk.cuh needs no arithmetic).  Undecidable-by-design cases
// (arithmetic beyond 1344 bits) answer "no": never a wrong yes.
//
// __host__ __device__ like the walker, so tests/hostsim can check it against repr(float(s)) on the CPU.
#ifndef CK_FLOAT_CUH
#define CK_FLOAT_CUH

#include <stdint.h>

#if defined(__CUDACC__)
#define CKF_HD __host__ __device__ __noinline__
#define CKF_IN __host__ __device__ __forceinline__
#else
#define CKF_HD __attribute__((noinline))
#define CKF_IN inline __attribute__((always_inline))
#endif

#define CKF_WORDS 42                      // 1344 bits

struct CkBig {                            // little-endian base 2^32, n significant words (n == 0: zero)
    uint32_t w[CKF_WORDS];
    uint32_t n;
    bool ovf;
};

CKF_IN void ckb_set(CkBig& a, uint64_t v) {
    a.ovf = false; a.n = 0;
    if (v) { a.w[a.n++] = (uint32_t)v; if (v >> 32) a.w[a.n++] = (uint32_t)(v >> 32); }
}
CKF_IN void ckb_mul_small(CkBig& a, uint32_t f) {
    uint64_t c = 0;
    for (uint32_t i = 0; i < a.n; i++) { uint64_t t = (uint64_t)a.w[i] * f + c; a.w[i] = (uint32_t)t; c = t >> 32; }
    if (c) { if (a.n < CKF_WORDS) a.w[a.n++] = (uint32_t)c; else a.ovf = true; }
}
CKF_IN void ckb_mul_pow10(CkBig& a, uint32_t p) {
    while (p >= 9) { ckb_mul_small(a, 1000000000u); p -= 9; }
    uint32_t f = 1;
    while (p--) f *= 10u;
    if (f != 1) ckb_mul_small(a, f);
}
CKF_IN void ckb_shl(CkBig& a, uint32_t bits) {
    if (a.n == 0 || bits == 0) return;
    uint32_t ws = bits >> 5, bs = bits & 31;
    if (a.n + ws + 1 > CKF_WORDS) { a.ovf = true; return; }
    if (bs) {
        uint32_t hi = a.w[a.n - 1] >> (32 - bs);
        for (uint32_t i = a.n - 1; i > 0; i--) a.w[i] = (a.w[i] << bs) | (a.w[i - 1] >> (32 - bs));
        a.w[0] <<= bs;
        if (hi) a.w[a.n++] = hi;
    }
    if (ws) {
        for (uint32_t i = a.n; i-- > 0;) a.w[i + ws] = a.w[i];
        for (uint32_t i = 0; i < ws; i++) a.w[i] = 0;
        a.n += ws;
    }
}
CKF_IN int ckb_cmp(const CkBig& a, const CkBig& b) {
    if (a.n != b.n) return a.n < b.n ? -1 : 1;
    for (uint32_t i = a.n; i-- > 0;) if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
    return 0;
}
CKF_IN uint32_t ckb_bits(const CkBig& a) {
    if (a.n == 0) return 0;
    uint32_t t = a.w[a.n - 1], b = 0;
    while (t) { b++; t >>= 1; }
    return (a.n - 1) * 32 + b;
}
// bits [lo, lo + 64) of a (lo may reach past the top: zeros)
CKF_IN uint64_t ckb_extract64(const CkBig& a, uint32_t lo) {
    uint64_t r = 0;
    for (uint32_t k = 0; k < 3; k++) {
        uint32_t wi = (lo >> 5) + k;
        if (wi >= a.n) break;
        uint64_t word = a.w[wi];
        int sh = (int)(32 * k) - (int)(lo & 31);
        if (sh >= 64) break;
        r |= sh >= 0 ? (word << sh) : (word >> (-sh));
    }
    return r;
}
CKF_IN bool ckb_any_below(const CkBig& a, uint32_t bit) {      // any set bit strictly below `bit`?
    uint32_t wi = bit >> 5;
    for (uint32_t i = 0; i < wi && i < a.n; i++) if (a.w[i]) return true;
    if (wi < a.n && (bit & 31)) return (a.w[wi] & ((1u << (bit & 31)) - 1u)) != 0;
    return false;
}

// sign of  a * 10^pa * 2^sa  -  b * 10^pb * 2^sb   (all exponents >= 0);  2 = arithmetic overflow (give up)
CKF_HD int ckf_cmp_scaled(uint64_t a, uint32_t pa, uint32_t sa, uint64_t b, uint32_t pb, uint32_t sb) {
    // cancel the common power of ten and two first: keeps the integers small in the common case
    uint32_t pc = pa < pb ? pa : pb, sc = sa < sb ? sa : sb;
    pa -= pc; pb -= pc; sa -= sc; sb -= sc;
    CkBig x, y;
    ckb_set(x, a); ckb_mul_pow10(x, pa); ckb_shl(x, sa);
    ckb_set(y, b); ckb_mul_pow10(y, pb); ckb_shl(y, sb);
    if (x.ovf || y.ovf) return 2;
    return ckb_cmp(x, y);
}

// decimal m * 10^k  vs  binary g * 2^t     -> -1 / 0 / +1, 2 = give up
CKF_IN int ckf_cmp_dec_bin(uint64_t m, int k, uint64_t g, int t) {
    // m*10^k ? g*2^t   <=>   m * 10^max(k,0) * 2^max(-t,0)  ?  g * 10^max(-k,0) * 2^max(t,0)
    return ckf_cmp_scaled(m, k > 0 ? (uint32_t)k : 0u, t < 0 ? (uint32_t)(-t) : 0u,
                          g, k < 0 ? (uint32_t)(-k) : 0u, t > 0 ? (uint32_t)t : 0u);
}

// d = f * 2^e nearest to m * 10^k (ties to even): 53-bit f for a normal double, f < 2^52 with e == -1074 for a subnormal.
// Returns CKF_FINITE, CKF_INF (rounds to infinity), CKF_ZERO (rounds to zero) or CKF_UNDECIDED (outside what the bignums hold).
enum { CKF_FINITE = 0, CKF_UNDECIDED = 1, CKF_INF = 2, CKF_ZERO = 3 };
CKF_HD int ckf_nearest_double_ex(uint64_t m, int k, uint64_t& f, int& e) {
    if (m == 0 || k > 310 || k < -345) return CKF_UNDECIDED;
    // work on the rational N / D with N = m * 10^max(k,0), D = 10^max(-k,0); scale N by 2^s so that the integer
    // quotient has 55..56 bits, then round to 53
    CkBig N, D;
    ckb_set(N, m); ckb_mul_pow10(N, k > 0 ? (uint32_t)k : 0u);
    ckb_set(D, 1); ckb_mul_pow10(D, k < 0 ? (uint32_t)(-k) : 0u);
    if (N.ovf || D.ovf) return CKF_UNDECIDED;
    int nb = (int)ckb_bits(N), db = (int)ckb_bits(D);
    int s = 56 - (nb - db);                       // quotient of (N << s) / D has 56 or 57 bits
    int e2;                                       // v = (N * 2^s / D) * 2^-s
    uint64_t q; bool rem;
    if (k >= 0) {
        // D == 1: the quotient is N itself, shifted to 56..57 bits
        if (s >= 0) { q = ckb_extract64(N, 0) << s; rem = false; }          // N has <= 57 bits here
        else { q = ckb_extract64(N, (uint32_t)(-s)); rem = ckb_any_below(N, (uint32_t)(-s)); }
        e2 = -s;
    } else {
        // long division with a 64-bit quotient: estimate from the leading bits, then correct exactly
        CkBig Ns = N;
        if (s > 0) ckb_shl(Ns, (uint32_t)s);
        CkBig Ds = D;
        if (s < 0) ckb_shl(Ds, (uint32_t)(-s));
        if (Ns.ovf || Ds.ovf) return CKF_UNDECIDED;
        uint32_t dbits = ckb_bits(Ds);
        uint32_t drop = dbits > 60 ? dbits - 60 : 0;          // leading 60 bits of the divisor
        uint64_t dtop = ckb_extract64(Ds, drop);
        // numerator's bits above `drop`: up to 60 + 57 bits -> take it as two 64-bit halves and divide
        uint64_t nlo = ckb_extract64(Ns, drop), nhi = ckb_extract64(Ns, drop + 64);
        unsigned __int128 num = ((unsigned __int128)nhi << 64) | nlo;
        uint64_t qe = (uint64_t)(num / dtop);
        // exact remainder sign via multiplication: find q with q*Ds <= Ns < (q+1)*Ds, |q - qe| is tiny
        q = qe;
        for (int iter = 0; iter < 8; iter++) {
            CkBig P = Ds;
            // P = Ds * q  (q < 2^58): two 32-bit multiplies
            CkBig Plo = Ds; ckb_mul_small(Plo, (uint32_t)q);
            CkBig Phi = Ds; ckb_mul_small(Phi, (uint32_t)(q >> 32)); ckb_shl(Phi, 32);
            // P = Plo + Phi
            uint64_t c = 0; uint32_t nn = Plo.n > Phi.n ? Plo.n : Phi.n;
            for (uint32_t i = 0; i < nn; i++) {
                uint64_t t = c + (i < Plo.n ? Plo.w[i] : 0u) + (uint64_t)(i < Phi.n ? Phi.w[i] : 0u);
                P.w[i] = (uint32_t)t; c = t >> 32;
            }
            P.n = nn; P.ovf = Plo.ovf || Phi.ovf;
            if (c) { if (P.n < CKF_WORDS) P.w[P.n++] = (uint32_t)c; else P.ovf = true; }
            while (P.n && P.w[P.n - 1] == 0) P.n--;
            if (P.ovf) return CKF_UNDECIDED;
            int c1 = ckb_cmp(P, Ns);
            if (c1 > 0) { q--; continue; }                     // q too large
            // P <= Ns: is Ns - P < Ds ?  <=>  P + Ds > Ns
            CkBig S = P; uint64_t cc = 0; uint32_t n2 = S.n > Ds.n ? S.n : Ds.n;
            for (uint32_t i = 0; i < n2; i++) {
                uint64_t t = cc + (i < S.n ? S.w[i] : 0u) + (uint64_t)(i < Ds.n ? Ds.w[i] : 0u);
                S.w[i] = (uint32_t)t; cc = t >> 32;
            }
            S.n = n2;
            if (cc) { if (S.n < CKF_WORDS) S.w[S.n++] = (uint32_t)cc; else return CKF_UNDECIDED; }
            if (ckb_cmp(S, Ns) <= 0) { q++; continue; }        // q too small
            rem = (c1 != 0);
            goto have_q;
        }
        return CKF_UNDECIDED;
have_q:
        e2 = -s;
    }
    // q has 56..57 bits (value = (q + rem_fraction) * 2^e2): round to 53 bits — or to the subnormal grid 2^-1074 when
    // that is coarser — ties to even
    int qb = 0; { uint64_t t = q; while (t) { qb++; t >>= 1; } }
    if (qb < 54) return CKF_UNDECIDED;
    int drop2 = qb - 53;
    if (e2 + drop2 < -1074) drop2 = -1074 - e2;              // subnormal: fewer than 53 bits survive
    if (drop2 > qb) return CKF_ZERO;                          // v < 2^qb * 2^e2 <= half of the grid step
    uint64_t keep = q >> drop2, low = q & ((1ull << drop2) - 1), half = 1ull << (drop2 - 1);      // 3 <= drop2 <= 57
    bool up;
    if (low > half) up = true;
    else if (low < half) up = false;
    else if (rem) up = true;
    else up = (keep & 1) != 0;                   // exact tie: to even
    f = keep + (up ? 1 : 0);
    e = e2 + drop2;
    if (f == 0) return CKF_ZERO;
    if (f == (1ull << 53)) { f >>= 1; e += 1; }
    int fb = 0; { uint64_t t = f; while (t) { fb++; t >>= 1; } }
    int exp2 = e + fb - 1;                        // value = 1.xxx * 2^exp2
    if (exp2 > 1023) return CKF_INF;
    if (exp2 < -1074) return CKF_ZERO;
    return CKF_FINITE;
}
CKF_HD bool ckf_nearest_double(uint64_t m, int k, uint64_t& f, int& e) { return ckf_nearest_double_ex(m, k, f, e) == CKF_FINITE; }

// m: the literal's significant digits as an integer (no trailing zeros, 16 or 17 digits), k: decimal exponent of
// its last digit.  true only if the literal is exactly repr() of the double it denotes.
CKF_HD bool ckf_is_repr(uint64_t m, int k) {
    uint64_t f; int e;
    if (!ckf_nearest_double(m, k, f, e)) return false;
    // rounding interval of d = f*2^e:  ( (2f-1)*2^(e-1) , (2f+1)*2^(e-1) ), endpoints included iff f even;
    // below a power of two the lower half-gap is half as wide: (4f-1)*2^(e-2)
    bool pow2 = (f == (1ull << 52)) && e > -1074;       // (below the smallest normal the spacing does not halve)
    uint64_t lo_g = pow2 ? 4 * f - 1 : 2 * f - 1; int lo_t = pow2 ? e - 2 : e - 1;
    uint64_t hi_g = 2 * f + 1; int hi_t = e - 1;
    bool incl = (f & 1) == 0;
    // does c * 10^kk round to d?  1 yes, 0 no, 2 give up
#define CKF_INSIDE(res, c, kk) do { int a_ = ckf_cmp_dec_bin((c), (kk), lo_g, lo_t), b_ = ckf_cmp_dec_bin((c), (kk), hi_g, hi_t); \
        (res) = (a_ == 2 || b_ == 2) ? 2 : (((a_ > 0 || (a_ == 0 && incl)) && (b_ < 0 || (b_ == 0 && incl))) ? 1 : 0); } while (0)
    // (B) the two (n-1)-digit neighbours c * 10^(k+1), c = floor(m/10), c+1, must fall outside the interval
    for (uint64_t c = m / 10; c <= m / 10 + 1; c++) {
        int in; CKF_INSIDE(in, c, k + 1);
        if (in != 0) return false;                               // a shorter spelling round-trips (or undecided)
    }
    // (C) among the n-digit decimals that round to d, repr()/ryu print the one nearest to d, an exact tie going to
    // the even digit.  v qualifies if it is strictly nearest, or if the grid point that beats (or ties) it does not
    // round to d — which happens just above a power of two, where the interval is half as wide on the low side.
    int c1 = ckf_cmp_dec_bin(2 * m - 1, k, f, e + 1);            // (2m-1)*10^k vs 2d:  < 0  <=>  v-1 is farther than v
    int c2 = ckf_cmp_dec_bin(2 * m + 1, k, f, e + 1);            // (2m+1)*10^k vs 2d:  > 0  <=>  v+1 is farther than v
    if (c1 == 2 || c2 == 2) return false;
    if (c1 < 0 && c2 > 0) return true;
    uint64_t rival = (c1 >= 0) ? m - 1 : m + 1;                  // the neighbour at least as close to d as v
    bool tie = (c1 == 0 || c2 == 0);
    if (tie && (m & 1) == 0) return true;
    int in; CKF_INSIDE(in, rival, k);
    return in == 0;
#undef CKF_INSIDE
}

// Shortest round-trip spelling of the double a literal denotes (what pydantic-core / repr() print for it), for literals that
// are NOT already that spelling: m * 10^k is the literal (m without trailing zeros, <= 19 digits), the result is ms * 10^ks
// (ms without trailing zeros).  Search instead of digit generation, every step decided with the exact comparisons above:
//   level n = 1 .. 17 digits: the n-digit decimals next to the literal (its truncation t and t + 1) are the only candidates
//   for "some n-digit decimal rounds to d" — the literal lies inside d's rounding interval, so any n-digit point inside it
//   has one of the two between itself and the literal; at the first level that has one, the answer is the grid point nearest
//   to d (found by stepping over midpoints) that rounds to d, confirmed by ckf_is_repr — which also arbitrates ties and the
//   narrow interval above a power of two — with its two neighbours as fallbacks.
// false = undecided (the caller reports CK_UNSUPPORTED): never a wrong spelling.
CKF_HD bool ckf_shortest(uint64_t m, int k, uint64_t& ms, int& ks) {
    uint64_t f; int e;
    if (!ckf_nearest_double(m, k, f, e)) return false;
    bool pow2 = (f == (1ull << 52)) && e > -1074;
    uint64_t lo_g = pow2 ? 4 * f - 1 : 2 * f - 1; int lo_t = pow2 ? e - 2 : e - 1;
    uint64_t hi_g = 2 * f + 1; int hi_t = e - 1;
    bool incl = (f & 1) == 0;
#define CKF_INSIDE2(res, c, kk) do { int a_ = ckf_cmp_dec_bin((c), (kk), lo_g, lo_t), b_ = ckf_cmp_dec_bin((c), (kk), hi_g, hi_t); \
        (res) = (a_ == 2 || b_ == 2) ? 2 : (((a_ > 0 || (a_ == 0 && incl)) && (b_ < 0 || (b_ == 0 && incl))) ? 1 : 0); } while (0)
    uint32_t nd = 0; { uint64_t t = m; while (t) { nd++; t /= 10; } }
    if (nd < 1 || nd > 19) return false;
    // a normal double needs at least 15 digits only if the literal has them: start where an answer can first exist
    for (uint32_t n = 1; n <= 17 && n <= nd; n++) {
        uint32_t p = nd - n;
        uint64_t pw = 1; for (uint32_t i = 0; i < p; i++) pw *= 10;
        uint64_t t = m / pw; int kn = k + (int)p;
        int in0 = 0, in1 = 0;
        if (t) CKF_INSIDE2(in0, t, kn);
        if (p > 0) CKF_INSIDE2(in1, t + 1, kn);
        if (in0 == 2 || in1 == 2) return false;
        if (!in0 && !in1) continue;
        uint64_t c = in0 ? t : t + 1;
        // nearest grid point to d: smallest c whose upper midpoint is not below d, then not above its lower midpoint
        for (int it = 0; it < 24; it++) { int g = ckf_cmp_dec_bin(2 * c + 1, kn, f, e + 1); if (g == 2) return false; if (g < 0) c++; else break; }
        for (int it = 0; it < 24 && c > 1; it++) { int g = ckf_cmp_dec_bin(2 * c - 1, kn, f, e + 1); if (g == 2) return false; if (g > 0) c--; else break; }
        uint64_t pick = 0; int pick_k = kn; bool found = false;
        for (int dlt = 0; dlt < 3 && !found; dlt++) {
            uint64_t cc = dlt == 0 ? c : (dlt == 1 ? c - 1 : c + 1);
            if (cc == 0) continue;
            int kc = kn;
            while (cc % 10 == 0) { cc /= 10; kc++; }          // a carry (9 -> 10): the same value with fewer digits
            uint32_t cd = 0; { uint64_t tt = cc; while (tt) { cd++; tt /= 10; } }
            if (cd > n) continue;
            int ins; CKF_INSIDE2(ins, cc, kc);                // it must denote d, not a neighbouring double
            if (ins == 2) return false;
            if (ins == 1 && ckf_is_repr(cc, kc)) { pick = cc; pick_k = kc; found = true; }
        }
        if (!found) return false;
        ms = pick; ks = pick_k;
        return true;
    }
    return false;
#undef CKF_INSIDE2
}

#endif
