// dump_parse.cpp -- host-side LAMMPS dump reader (replaces pymatgen.io.lammps.outputs.parse_lammps_dumps,
// called at mdproptools/structural/rdf_cn.py:176, cluster_analysis.py:100, hydration_number.py:84,
// dynamical/diffusion.py:172, conductivity.py:87, residence_time.py:54).
//
// Text format per frame:
//   ITEM: TIMESTEP / <int> / ITEM: NUMBER OF ATOMS / <int> / ITEM: BOX BOUNDS [xy xz yz] pp pp pp /
//   3 bound lines (lo hi [tilt]) / ITEM: ATOMS <column names> / natoms rows
// Numbers are converted with std::from_chars (correctly rounded, == Python float() == pandas' parser on
// this data), rows are scattered to id-1 (the reference sorts every frame by id, rdf_cn.py:191-192), and
// the bounds get the tilt correction pymatgen applies before building the box.
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <memory>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mdprop_b200.h"
#include "dump_line.h"

void mdp_set_error(const char *fmt, ...);

namespace {

struct Header {
    long long timestep = 0, natoms = 0;
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, tilt[3] = {0, 0, 0};
    bool triclinic = false;
    std::vector<std::string> cols;
    const char *atoms_begin = nullptr;   // first byte after the "ITEM: ATOMS ..." line
};

inline const char *next_line(const char *p, const char *end)
{
    const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
    return nl ? nl + 1 : end;
}

inline const char *skip_ws(const char *p, const char *end) { return mdp_skip_ws(p, end); }
inline const char *token_end(const char *p, const char *end) { return mdp_token_end(p, end); }

// Decimal -> double.  Fast path (Clinger): a decimal significand w <= 2^53 and a power of ten |q| <= 22 are both exact
// doubles, so ONE correctly rounded IEEE multiplication or division gives the correctly rounded result -- the same
// double std::from_chars / strtod / Python float() return.  LAMMPS writes %g (6 significant digits) unless told
// otherwise, so practically every token of a dump takes this path; everything else (more digits, big exponents,
// inf/nan, malformed text) goes to std::from_chars.

inline bool parse_double(const char *b, const char *e, double &v)
{
    if (b < e && *b == '+') ++b;
    const char *p = b;
    bool neg = false;
    if (p < e && *p == '-') {
        neg = true;
        ++p;
    }
    uint64_t w = 0;
    int nd = 0, q = 0;          // significant digits taken into w, decimal exponent
    bool any = false, ok = true;
    while (p < e && (unsigned)(*p - '0') < 10u) {
        any = true;
        if (w || *p != '0') {
            if (nd < 19) {
                w = w * 10 + (unsigned)(*p - '0');
                ++nd;
            } else {
                ok = false;
            }
        }
        ++p;
    }
    if (p < e && *p == '.') {
        ++p;
        while (p < e && (unsigned)(*p - '0') < 10u) {
            any = true;
            if (w || *p != '0') {
                if (nd < 19) {
                    w = w * 10 + (unsigned)(*p - '0');
                    ++nd;
                } else {
                    ok = false;
                }
            }
            --q;
            ++p;
        }
    }
    if (any && ok && p < e && (*p == 'e' || *p == 'E')) {
        const char *r = p + 1;
        bool eneg = false;
        if (r < e && (*r == '-' || *r == '+')) {
            eneg = *r == '-';
            ++r;
        }
        int ex = 0, ne = 0;
        while (r < e && (unsigned)(*r - '0') < 10u && ne < 6) {
            ex = ex * 10 + (*r - '0');
            ++ne;
            ++r;
        }
        if (ne == 0 || ne >= 6) ok = false;
        q += eneg ? -ex : ex;
        p = r;
    }
    if (any && ok && p == e && w <= (1ull << 53) && q >= -22 && q <= 22) {
        double d = (double)w;
        d = q < 0 ? d / mdp_pow10(-q) : d * mdp_pow10(q);
        v = neg ? -d : d;
        return true;
    }
    auto r = std::from_chars(b, e, v);
    return r.ec == std::errc() && r.ptr == e;
}

int parse_header(const char *text, int64_t len, Header &h)
{
    const char *p = text, *end = text + len;
    if (len < 14 || strncmp(p, "ITEM: TIMESTEP", 14) != 0) {
        mdp_set_error("dump: frame does not start with 'ITEM: TIMESTEP'");
        return -4;
    }
    p = next_line(p, end);
    h.timestep = strtoll(p, nullptr, 10);
    p = next_line(p, end);   // ITEM: NUMBER OF ATOMS
    p = next_line(p, end);
    h.natoms = strtoll(p, nullptr, 10);
    p = next_line(p, end);   // ITEM: BOX BOUNDS ...
    {
        const char *le = next_line(p, end);
        std::string line(p, le);
        h.triclinic = line.find("xy xz yz") != std::string::npos;
        p = le;
    }
    for (int k = 0; k < 3; ++k) {
        const char *le = next_line(p, end);
        const char *q = skip_ws(p, le);
        double vals[3] = {0, 0, 0};
        int nv = 0;
        while (q < le && *q != '\n' && nv < 3) {
            const char *te = token_end(q, le);
            if (te == q) break;
            if (!parse_double(q, te, vals[nv])) {
                mdp_set_error("dump: cannot parse box bound '%.*s'", (int)(te - q), q);
                return -4;
            }
            ++nv;
            q = skip_ws(te, le);
        }
        if (nv < 2) {
            mdp_set_error("dump: box bound line %d has %d values", k, nv);
            return -4;
        }
        h.lo[k] = vals[0];
        h.hi[k] = vals[1];
        h.tilt[k] = nv > 2 ? vals[2] : 0.0;
        p = le;
    }
    if (h.triclinic) {
        // pymatgen LammpsDump.from_string: bounds -= [[min(0,xy,xz,xy+xz), max(...)], [min(0,yz), max(0,yz)], [0,0]]
        const double xy = h.tilt[0], xz = h.tilt[1], yz = h.tilt[2];
        const double xs[4] = {0.0, xy, xz, xy + xz};
        h.lo[0] -= *std::min_element(xs, xs + 4);
        h.hi[0] -= *std::max_element(xs, xs + 4);
        h.lo[1] -= std::min(0.0, yz);
        h.hi[1] -= std::max(0.0, yz);
    }
    {
        const char *le = next_line(p, end);
        if (le - p < 11 || strncmp(p, "ITEM: ATOMS", 11) != 0) {
            mdp_set_error("dump: expected 'ITEM: ATOMS' line");
            return -4;
        }
        const char *q = p + 11;
        while (true) {
            q = skip_ws(q, le);
            if (q >= le || *q == '\n') break;
            const char *te = token_end(q, le);
            h.cols.emplace_back(q, te);
            q = te;
        }
        p = le;
    }
    h.atoms_begin = p;
    return 0;
}

void fill_header_out(const Header &h, double *o, bool id_contiguous)
{
    o[0] = (double)h.timestep;
    o[1] = (double)h.natoms;
    for (int k = 0; k < 3; ++k) {
        o[2 + 2 * k] = h.lo[k];
        o[3 + 2 * k] = h.hi[k];
    }
    o[8] = h.tilt[0];
    o[9] = h.tilt[1];
    o[10] = h.tilt[2];
    o[11] = h.triclinic ? 1.0 : 0.0;
    o[12] = (double)h.cols.size();
    o[13] = id_contiguous ? 1.0 : 0.0;
    o[14] = 0.0;
    o[15] = 0.0;
}

// parse rows [row0, row0+nrows) starting at byte p; colsel[c] = output slot of file column c or -1
int parse_rows(const char *p, const char *end, int64_t row0, int64_t nrows, const std::vector<int> &colsel, int id_col,
               double *vals /* [nslots][natoms] file order */, int64_t natoms, long long *ids)
{
    const int ncols = (int)colsel.size();
    for (int64_t r = 0; r < nrows; ++r) {
        const char *le = next_line(p, end);
        while (p < end && skip_ws(p, le) >= le - (le > p && le[-1] == '\n' ? 1 : 0)) {   // blank line (pandas skips them too)
            p = le;
            le = next_line(p, end);
        }
        const char *q = p;
        for (int c = 0; c < ncols; ++c) {
            q = skip_ws(q, le);
            const char *te = token_end(q, le);
            if (te == q) {
                mdp_set_error("dump: row %lld has fewer than %d columns", (long long)(row0 + r), ncols);
                return -4;
            }
            if (c == id_col) {
                long long idv = 0;
                auto rr = std::from_chars(q, te, idv);
                if (rr.ec != std::errc()) {
                    double dv;
                    if (!parse_double(q, te, dv)) {
                        mdp_set_error("dump: bad id '%.*s'", (int)(te - q), q);
                        return -4;
                    }
                    idv = (long long)dv;
                }
                ids[row0 + r] = idv;
            }
            const int slot = colsel[c];
            if (slot >= 0) {
                double v;
                if (!parse_double(q, te, v)) {
                    mdp_set_error("dump: cannot parse '%.*s' in row %lld column %d", (int)(te - q), q, (long long)(row0 + r), c);
                    return -4;
                }
                vals[(int64_t)slot * natoms + row0 + r] = v;
            }
            q = te;
        }
        p = le;
    }
    return 0;
}

// Token at q (no leading blanks): the exact one-pass fast path shared with the device parser (dump_line.h), else the
// token is re-scanned and handed to parse_double.  Returns the end of the token, nullptr when it is empty or malformed.
inline const char *parse_token(const char *q, const char *le, double &v)
{
    if (const char *p = mdp_parse_fast(q, le, &v)) return p;
    const char *te = token_end(q, le);
    if (te == q || !parse_double(q, te, v)) return nullptr;
    return te;
}

// column selection shared by the single-frame and the batch entry points
int select_columns(const Header &h, const char *const *want, int nwant, std::vector<int> &colsel, int &id_col)
{
    const int ncols = (int)h.cols.size();
    colsel.assign(ncols, -1);
    id_col = -1;
    for (int c = 0; c < ncols; ++c)
        if (h.cols[c] == "id") id_col = c;
    for (int k = 0; k < nwant; ++k) {
        int found = -1;
        for (int c = 0; c < ncols; ++c)
            if (h.cols[c] == want[k]) found = c;
        if (found < 0) {
            mdp_set_error("dump: column '%s' not present in the dump file", want[k]);
            return -4;
        }
        if (colsel[found] >= 0) {
            mdp_set_error("mdp_dump_parse: column '%s' requested twice", want[k]);
            return -2;
        }
        colsel[found] = k;
    }
    return 0;
}

// Rows written straight to their id-sorted place (out[slot][id - 1]) while the ids are a permutation of 1..natoms -- the
// usual case.  Parses the first `nrows` non-blank rows of [p, end); `seen` is a bitmap of natoms bits shared by every
// thread working on the frame (ATOMIC) or private to the caller.  Returns 0 on success, 1 when an id is out of range or
// occurs twice (the caller then takes the general path, which ranks the ids), < 0 on error.
template <bool ATOMIC>
int parse_rows_fused(const char *p, const char *end, int64_t row0, int64_t nrows, int64_t n, const std::vector<int> &colsel,
                     int id_col, int nwant, double *out, int64_t out_stride, std::atomic<uint32_t> *seen)
{
    if (id_col < 0 || nwant > 64) return 1;
    const int ncols = (int)colsel.size();
    int last_needed = id_col;
    for (int c = 0; c < ncols; ++c)
        if (colsel[c] >= 0 && c > last_needed) last_needed = c;
    double rowv[64];
    int64_t r = 0;
    while (r < nrows) {
        if (p >= end) {
            mdp_set_error("dump: frame announces %lld atoms but holds %lld rows", (long long)n, (long long)(row0 + r));
            return -4;
        }
        const char *le = next_line(p, end);
        const char *q = skip_ws(p, le);
        if (q >= le || *q == '\n') {   // blank line
            p = le;
            continue;
        }
        long long idv = 0;
        for (int c = 0; c <= last_needed; ++c) {
            q = skip_ws(q, le);
            const int slot = colsel[c];
            if (slot < 0 && c != id_col) {          // unwanted column: only find its end
                const char *te = token_end(q, le);
                if (te == q) {
                    mdp_set_error("dump: row %lld has fewer than %d columns", (long long)(row0 + r), ncols);
                    return -4;
                }
                q = te;
                continue;
            }
            double v;
            const char *te = parse_token(q, le, v);
            if (!te) {
                const char *t2 = token_end(q, le);
                if (t2 == q)
                    mdp_set_error("dump: row %lld has fewer than %d columns", (long long)(row0 + r), ncols);
                else
                    mdp_set_error("dump: cannot parse '%.*s' in row %lld column %d", (int)(t2 - q), q, (long long)(row0 + r), c);
                return -4;
            }
            if (c == id_col) {
                // ids are integers; a spelling like 1e3 or 12.0 goes through the double
                if (!(v >= -9.2e18 && v <= 9.2e18)) return 1;
                idv = (long long)v;
            }
            if (slot >= 0) rowv[slot] = v;
            q = te;
        }
        if (idv < 1 || idv > n) return 1;
        const uint32_t bit = 1u << ((idv - 1) & 31);
        std::atomic<uint32_t> &word = seen[(idv - 1) >> 5];
        if (ATOMIC) {
            if (word.fetch_or(bit, std::memory_order_relaxed) & bit) return 1;
        } else {
            const uint32_t o = word.load(std::memory_order_relaxed);
            if (o & bit) return 1;
            word.store(o | bit, std::memory_order_relaxed);
        }
        for (int k = 0; k < nwant; ++k) out[(int64_t)k * out_stride + (idv - 1)] = rowv[k];
        ++r;
        p = le;
    }
    return 0;
}

struct SeenBits {
    std::unique_ptr<std::atomic<uint32_t>[]> w;
    size_t words = 0;
    std::atomic<uint32_t> *reset(int64_t n)
    {
        const size_t need = (size_t)(n + 31) / 32;
        if (need > words) {
            w.reset(new std::atomic<uint32_t>[need]);
            words = need;
        }
        for (size_t k = 0; k < need; ++k) w[k].store(0, std::memory_order_relaxed);
        return w.get();
    }
};

// one frame on the calling thread
int parse_frame_fused(const Header &h, const char *end, const std::vector<int> &colsel, int id_col, int nwant, double *out,
                      int64_t out_stride, SeenBits &seen)
{
    return parse_rows_fused<false>(h.atoms_begin, end, 0, h.natoms, h.natoms, colsel, id_col, nwant, out, out_stride,
                                   seen.reset(h.natoms));
}

} // namespace

extern "C" {

int64_t mdp_dump_scan(const char *text, int64_t len, int64_t *offsets, int64_t max_frames)
{
    if (!text || len <= 0) return 0;
    // A frame starts at a line that begins with "ITEM: TIMESTEP".  Looking for the letter 'I' with memchr runs at memory
    // speed over the atom rows (numbers; a row that does hold an 'I', say an element column, only costs a comparison) --
    // a memchr per 30-byte line, as in round 1, cost three times the read of the file itself.
    int64_t n = 0;
    const char *end = text + len;
    for (const char *p = text; p < end;) {
        const char *q = (const char *)memchr(p, 'I', (size_t)(end - p));
        if (!q) break;
        if ((q == text || q[-1] == '\n') && end - q >= 14 && memcmp(q, "ITEM: TIMESTEP", 14) == 0) {
            if (offsets && n < max_frames) offsets[n] = (int64_t)(q - text);
            ++n;
            p = q + 14;
        } else {
            p = q + 1;
        }
    }
    return n;
}

int mdp_dump_header(const char *text, int64_t len, double *header_out, char *columns_out, int columns_cap)
{
    if (!text || !header_out) {
        mdp_set_error("mdp_dump_header: NULL argument");
        return -2;
    }
    Header h;
    int rc = parse_header(text, len, h);
    if (rc) return rc;
    fill_header_out(h, header_out, false);
    if (columns_out && columns_cap > 0) {
        std::string s;
        for (size_t k = 0; k < h.cols.size(); ++k) {
            if (k) s += ' ';
            s += h.cols[k];
        }
        if ((int)s.size() + 1 > columns_cap) {
            mdp_set_error("mdp_dump_header: column buffer too small (%zu needed)", s.size() + 1);
            return -2;
        }
        memcpy(columns_out, s.c_str(), s.size() + 1);
    }
    return 0;
}

int mdp_dump_parse(const char *text, int64_t len, const char *const *want, int nwant, double *out, int64_t out_stride,
                   double *header_out, int nthreads)
{
    if (!text || !want || !out || nwant <= 0) {
        mdp_set_error("mdp_dump_parse: bad argument");
        return -2;
    }
    Header h;
    int rc = parse_header(text, len, h);
    if (rc) return rc;
    const int64_t n = h.natoms;
    if (out_stride < n) {
        mdp_set_error("mdp_dump_parse: out_stride %lld < natoms %lld", (long long)out_stride, (long long)n);
        return -2;
    }
    std::vector<int> colsel;
    int id_col = -1;
    rc = select_columns(h, want, nwant, colsel, id_col);
    if (rc) return rc;
    const int ncols = (int)colsel.size();
    (void)ncols;

    const char *begin = h.atoms_begin, *end = text + len;
    // stop at the next frame if the buffer holds more than one
    // (rows never start with 'I'; the scan is a memchr per line only when needed)
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    nthreads = (int)std::min<int64_t>(nthreads, std::max<int64_t>(1, n / 4096));

    // split [begin, ...) into nthreads byte ranges on line boundaries, count rows per range
    std::vector<const char *> cut(nthreads + 1);
    cut[0] = begin;
    const int64_t approx = (int64_t)(end - begin);
    for (int t = 1; t < nthreads; ++t) {
        const char *g = begin + approx * t / nthreads;
        cut[t] = g <= cut[t - 1] ? cut[t - 1] : next_line(g, end);
    }
    cut[nthreads] = end;
    std::vector<int64_t> rows(nthreads, 0), row0(nthreads + 1, 0);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                int64_t c = 0;
                const char *p = cut[t];
                while (p < cut[t + 1]) {
                    const char *nl = (const char *)memchr(p, '\n', (size_t)(cut[t + 1] - p));
                    if (!nl) {
                        if (skip_ws(p, cut[t + 1]) < cut[t + 1]) ++c;   // last line without '\n'
                        break;
                    }
                    if (skip_ws(p, nl) < nl) ++c;
                    p = nl + 1;
                }
                rows[t] = c;
            });
        for (auto &x : th) x.join();
    }
    for (int t = 0; t < nthreads; ++t) row0[t + 1] = row0[t] + rows[t];
    if (row0[nthreads] < n) {
        mdp_set_error("dump: frame announces %lld atoms but holds %lld rows", (long long)n, (long long)row0[nthreads]);
        return -4;
    }

    std::vector<int> rcs(nthreads, 0);
    std::vector<std::string> errs(nthreads);
    if (id_col >= 0) {
        // usual case, ids a permutation of 1..natoms: every thread places its rows itself (shared atomic bitmap); no
        // staging copy, no serial pass
        SeenBits seen;
        std::atomic<uint32_t> *bits = seen.reset(n);
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                const int64_t r0 = row0[t];
                const int64_t nr = std::min<int64_t>(rows[t], std::max<int64_t>(0, n - r0));
                if (nr <= 0) return;
                rcs[t] = parse_rows_fused<true>(cut[t], cut[t + 1], r0, nr, n, colsel, id_col, nwant, out, out_stride, bits);
                if (rcs[t] < 0) errs[t] = mdp_last_error();
            });
        for (auto &x : th) x.join();
        bool irregular = false;
        for (int t = 0; t < nthreads; ++t) {
            if (rcs[t] < 0) {
                mdp_set_error("%s", errs[t].c_str());
                return rcs[t];
            }
            irregular = irregular || rcs[t] == 1;
        }
        if (!irregular) {
            if (header_out) fill_header_out(h, header_out, true);
            return 0;
        }
        std::fill(rcs.begin(), rcs.end(), 0);
    }

    // general path: values in file order, then ranked by id
    std::vector<double> vals((size_t)nwant * n);
    std::vector<long long> ids((size_t)n, 0);
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                const int64_t r0 = row0[t];
                const int64_t nr = std::min<int64_t>(rows[t], std::max<int64_t>(0, n - r0));
                if (nr <= 0) return;
                rcs[t] = parse_rows(cut[t], cut[t + 1], r0, nr, colsel, id_col, vals.data(), n, ids.data());
                if (rcs[t]) errs[t] = mdp_last_error();
            });
        for (auto &x : th) x.join();
    }
    for (int t = 0; t < nthreads; ++t)
        if (rcs[t]) {
            mdp_set_error("%s", errs[t].c_str());
            return rcs[t];
        }

    // order rows by id (reference: DataFrame.sort_values("id"))
    bool contiguous = id_col >= 0;
    std::vector<int64_t> dest((size_t)n);
    if (id_col >= 0) {
        std::vector<unsigned char> seen((size_t)n, 0);
        for (int64_t r = 0; r < n && contiguous; ++r) {
            const long long v = ids[r];
            if (v < 1 || v > n || seen[v - 1])
                contiguous = false;
            else {
                seen[v - 1] = 1;
                dest[r] = v - 1;
            }
        }
        if (!contiguous) {
            std::vector<int64_t> order((size_t)n);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return ids[a] < ids[b]; });
            for (int64_t k = 0; k < n; ++k) dest[order[k]] = k;
        }
    } else {
        std::iota(dest.begin(), dest.end(), 0);
    }
    for (int k = 0; k < nwant; ++k) {
        const double *src = vals.data() + (size_t)k * n;
        double *dst = out + (size_t)k * out_stride;
        for (int64_t r = 0; r < n; ++r) dst[dest[r]] = src[r];
    }
    if (header_out) fill_header_out(h, header_out, contiguous);
    return 0;
}

int mdp_dump_parse_batch(int nframes, const char *const *texts, const int64_t *lens, const char *const *want, int nwant,
                         double *out, int64_t frame_stride, int64_t out_stride, double *headers_out, int nthreads)
{
    if (nframes <= 0 || !texts || !lens || !want || !out || nwant <= 0) {
        mdp_set_error("mdp_dump_parse_batch: bad argument");
        return -2;
    }
    if (nthreads <= 0) nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    // few frames: every frame uses all threads (rows split over threads); many frames: one frame per thread at a time
    if (nframes * 2 < nthreads || nthreads == 1) {
        for (int f = 0; f < nframes; ++f) {
            int rc = mdp_dump_parse(texts[f], lens[f], want, nwant, out + (size_t)f * frame_stride, out_stride,
                                    headers_out ? headers_out + (size_t)f * 16 : nullptr, nthreads);
            if (rc) return rc;
        }
        return 0;
    }
    const int nt = std::min(nthreads, nframes);
    std::atomic<int> next(0), failed(0);
    std::vector<int> rcs(nt, 0);
    std::vector<std::string> errs(nt);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t]() {
            SeenBits seen;
            std::vector<int> colsel;
            while (!failed.load(std::memory_order_relaxed)) {
                const int f = next.fetch_add(1);
                if (f >= nframes) break;
                double *o = out + (size_t)f * frame_stride;
                double *ho = headers_out ? headers_out + (size_t)f * 16 : nullptr;
                Header h;
                int rc = parse_header(texts[f], lens[f], h);
                int id_col = -1;
                if (!rc && out_stride < h.natoms) {
                    mdp_set_error("mdp_dump_parse_batch: out_stride %lld < natoms %lld", (long long)out_stride, (long long)h.natoms);
                    rc = -2;
                }
                if (!rc) rc = select_columns(h, want, nwant, colsel, id_col);
                if (!rc) {
                    rc = parse_frame_fused(h, texts[f] + lens[f], colsel, id_col, nwant, o, out_stride, seen);
                    if (rc == 0 && ho) fill_header_out(h, ho, true);
                    if (rc == 1) rc = mdp_dump_parse(texts[f], lens[f], want, nwant, o, out_stride, ho, 1);   // general path
                }
                if (rc) {
                    rcs[t] = rc;
                    errs[t] = mdp_last_error();
                    failed.store(1);
                    break;
                }
            }
        });
    for (auto &x : th) x.join();
    for (int t = 0; t < nt; ++t)
        if (rcs[t]) {
            mdp_set_error("%s", errs[t].c_str());
            return rcs[t];
        }
    return 0;
}

} // extern "C"
