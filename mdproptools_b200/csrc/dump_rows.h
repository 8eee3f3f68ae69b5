// dump_rows.h -- the body of the device dump parser (dump_device.cu) as a host/device function, so that the very same
// code is exercised on the host by tests/native/dump_rows_host.cpp (every (frame, chunk) in turn, plain memory
// operations instead of atomics) and on the GPU by k_dump_rows (one thread per chunk, CUDA atomics).
#pragma once
#include "dump_line.h"

constexpr int DP_CHUNK = 64;       // bytes of text per thread
constexpr int DP_MAX_COLS = 64;    // columns of the file that are looked at (up to the last needed one)
constexpr int DP_MAX_WANT = 16;

enum { DPF_SLOW_TOKEN = 1, DPF_BAD_ROW = 2, DPF_BAD_ID = 4 };

struct DevParseParams {
    const char *text;                  // all frames' bytes
    const long long *begin, *end;      // per frame: first byte of the first row, one past the last byte of the last row
    long long natoms;
    int nlook;                         // columns 0..nlook-1 are tokenised
    int id_col, nwant;
    signed char colsel[DP_MAX_COLS];   // output slot of file column c, or -1
    double *out;                       // [F][nwant][out_stride]
    long long frame_stride, out_stride;
    unsigned *seen;                    // [F][seen_words]
    long long seen_words;
    unsigned long long *status;        // [F][2]: rows parsed, OR of DPF_* flags
};

// Chunk `chunk` of frame f = bytes [begin + chunk*DP_CHUNK, +DP_CHUNK) of the frame's row section: parse every row that
// STARTS in it (reading past the chunk's end as far as the row goes) and place it by id.
template <class Atomics>
MDP_HD void mdp_parse_chunk(const DevParseParams &p, int f, long long chunk, const Atomics at)
{
    const char *t = p.text;
    const long long b = p.begin[f], e = p.end[f];
    const long long c0 = b + chunk * DP_CHUNK;
    if (c0 >= e) return;
    const long long c1 = c0 + DP_CHUNK < e ? c0 + DP_CHUNK : e;
    // first row start inside [c0, c1)
    long long pos = c0;
    if (c0 > b && t[c0 - 1] != '\n') {
        while (pos < c1 && t[pos] != '\n') ++pos;
        ++pos;                          // one past the newline (>= c1 when the chunk holds no row start)
    }
    unsigned flags = 0;
    unsigned long long rows = 0;
    double *out = p.out + (long long)f * p.frame_stride;
    unsigned *seen = p.seen + (long long)f * p.seen_words;
    while (pos < c1) {
        long long le = pos;
        while (le < e && t[le] != '\n') ++le;
        const char *lend = t + le;
        const char *q = mdp_skip_ws(t + pos, lend);
        if (q < lend) {                 // not a blank line
            double rowv[DP_MAX_WANT];
            double idv = 0.0;
            bool ok = true;
            for (int c = 0; c < p.nlook; ++c) {
                q = mdp_skip_ws(q, lend);
                if (q >= lend) {
                    flags |= DPF_BAD_ROW;    // fewer columns than the header announces
                    ok = false;
                    break;
                }
                const int slot = p.colsel[c];
                if (slot < 0 && c != p.id_col) {
                    q = mdp_token_end(q, lend);
                    continue;
                }
                double v;
                const char *te = mdp_parse_fast(q, lend, &v);
                if (!te) {
                    flags |= DPF_SLOW_TOKEN;   // too many digits, scale beyond 10^+-22, inf/nan, malformed: the host parser decides
                    ok = false;
                    break;
                }
                if (c == p.id_col) idv = v;
                if (slot >= 0) rowv[slot] = v;
                q = te;
            }
            if (ok) {
                ++rows;
                if (idv >= 1.0 && idv <= (double)p.natoms) {
                    const long long id = (long long)idv - 1;
                    const unsigned bit = 1u << (id & 31);
                    if (at.or32(&seen[id >> 5], bit) & bit) flags |= DPF_BAD_ID;   // the id occurred before
                    for (int k = 0; k < p.nwant; ++k) out[(long long)k * p.out_stride + id] = rowv[k];
                } else {
                    flags |= DPF_BAD_ID;
                }
            }
        }
        pos = le + 1;
    }
    if (rows) at.add64(&p.status[f * 2 + 0], rows);
    if (flags) at.or64(&p.status[f * 2 + 1], (unsigned long long)flags);
}
