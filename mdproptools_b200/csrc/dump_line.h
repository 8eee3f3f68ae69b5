// dump_line.h -- the exact fast path of the LAMMPS dump row parser, shared by the host parser (dump_parse.cpp) and the
// device parser (dump_device.cu).  Everything here is plain integer / IEEE fp64 arithmetic that gives the same bits on
// the host and on the GPU.
//
// Decimal -> double (Clinger): a decimal significand w <= 2^53 and a power of ten 10^k, k <= 22, are both exact doubles,
// so ONE correctly rounded IEEE division (or multiplication) gives the correctly rounded result -- the double that
// std::from_chars / strtod / Python float() / pandas return for the same text.  LAMMPS writes %g (6 significant digits)
// unless told otherwise, so practically every token of a dump qualifies; whatever does not (more digits, exponents
// beyond +-22, inf/nan, malformed text) is left to the caller's exact general parser.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MDP_HD __host__ __device__ __forceinline__
#else
#define MDP_HD inline
#endif

MDP_HD double mdp_pow10(int k)   // 0 <= k <= 22: exact
{
    constexpr double t[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                              1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    return t[k];
}

MDP_HD bool mdp_is_blank(char c) { return c == ' ' || c == '\t' || c == '\r'; }

MDP_HD const char *mdp_skip_ws(const char *p, const char *end)
{
    while (p < end && mdp_is_blank(*p)) ++p;
    return p;
}

MDP_HD const char *mdp_token_end(const char *p, const char *end)
{
    while (p < end && !mdp_is_blank(*p) && *p != '\n') ++p;
    return p;
}

// Token at q (no leading blanks), scanned and converted in one pass.  Returns the end of the token when the exact fast
// path applies (v holds the correctly rounded value), nullptr otherwise (v untouched): the token then has too many
// digits, a decimal scale beyond 10^+-22, no digit at all, or does not end at a blank / line end.
MDP_HD const char *mdp_parse_fast(const char *q, const char *le, double *v)
{
    const char *p = q;
    bool neg = false;
    if (p < le && (*p == '-' || *p == '+')) {
        neg = *p == '-';
        ++p;
    }
    uint64_t w = 0;
    int nd = 0, sc = 0;      // digits from the first non-zero one on; minus the number of fraction digits
    const char *d0 = p;
    while (p < le && (unsigned)(*p - '0') < 10u) {
        w = w * 10 + (unsigned)(*p - '0');
        nd += (w != 0);
        ++p;
    }
    bool any = p > d0;
    if (p < le && *p == '.') {
        ++p;
        const char *f0 = p;
        while (p < le && (unsigned)(*p - '0') < 10u) {
            w = w * 10 + (unsigned)(*p - '0');
            nd += (w != 0);
            ++p;
        }
        sc = (int)(f0 - p);
        any = any || p > f0;
    }
    if (any && p < le && (*p == 'e' || *p == 'E')) {     // decimal exponent (%g writes one below 1e-4: "1.23e-05")
        const char *x = p + 1;
        bool xneg = false;
        if (x < le && (*x == '-' || *x == '+')) {
            xneg = *x == '-';
            ++x;
        }
        const char *x0 = x;
        int ex = 0;
        while (x < le && (unsigned)(*x - '0') < 10u && x - x0 < 4) {
            ex = ex * 10 + (*x - '0');
            ++x;
        }
        if (x == x0) return nullptr;                     // "1e" / "1e+": not a number the fast path knows
        sc += xneg ? -ex : ex;
        p = x;
    }
    // w cannot have wrapped while nd <= 19 (10^19 - 1 < 2^64).  Exact (Clinger): w and 10^|sc| are both doubles, so the
    // one division or multiplication is the only rounding
    if (any && nd <= 19 && w <= (1ull << 53) && sc >= -22 && sc <= 22 && (p == le || mdp_is_blank(*p) || *p == '\n')) {
        double d = (double)w;
        if (sc < 0) d = d / mdp_pow10(-sc);
        else if (sc > 0) d = d * mdp_pow10(sc);
        *v = neg ? -d : d;
        return p;
    }
    return nullptr;
}
