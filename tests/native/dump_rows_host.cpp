// Host emulation of k_dump_rows (mdproptools_b200/csrc/dump_device.cu): the same mdp_parse_chunk the kernel runs,
// called for every (frame, chunk) in turn with plain memory operations.  TEST INFRASTRUCTURE ONLY: built by
// tests/test_host_logic.py with g++ and loaded through ctypes.
#include <stdint.h>
#include <string.h>

#include "../../mdproptools_b200/csrc/dump_rows.h"

struct HostOps {
    unsigned or32(unsigned *p, unsigned v) const
    {
        const unsigned o = *p;
        *p = o | v;
        return o;
    }
    void add64(unsigned long long *p, unsigned long long v) const { *p += v; }
    void or64(unsigned long long *p, unsigned long long v) const { *p |= v; }
};

extern "C" int emulate_dump_rows(int nframes, const char *text, const long long *begin, const long long *end, long long natoms,
                                 int ncols, const int *colsel, int id_col, int nwant, double *out, long long frame_stride,
                                 long long out_stride, unsigned *seen, unsigned long long *status, int reverse_order)
{
    DevParseParams p;
    memset(&p, 0, sizeof(p));
    int last = id_col;
    for (int c = 0; c < ncols; ++c)
        if (colsel[c] >= 0 && c > last) last = c;
    if (last >= DP_MAX_COLS || nwant > DP_MAX_WANT) return -1;
    p.nlook = last + 1;
    for (int c = 0; c < DP_MAX_COLS; ++c) p.colsel[c] = (signed char)(c < ncols && c < p.nlook ? colsel[c] : -1);
    p.text = text;
    p.begin = begin;
    p.end = end;
    p.natoms = natoms;
    p.id_col = id_col;
    p.nwant = nwant;
    p.out = out;
    p.frame_stride = frame_stride;
    p.out_stride = out_stride;
    p.seen = seen;
    p.seen_words = (natoms + 31) / 32;
    p.status = status;
    memset(seen, 0, (size_t)nframes * p.seen_words * 4);
    memset(status, 0, (size_t)nframes * 16);
    for (int f = 0; f < nframes; ++f) {
        const long long chunks = (end[f] - begin[f] + DP_CHUNK - 1) / DP_CHUNK;
        // threads run in no particular order on the GPU: both directions must give the same result
        for (long long k = 0; k < chunks + 3; ++k) mdp_parse_chunk(p, f, reverse_order ? chunks + 2 - k : k, HostOps());
    }
    return 0;
}
