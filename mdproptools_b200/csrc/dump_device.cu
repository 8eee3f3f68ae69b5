// dump_device.cu -- device-side parser of LAMMPS dump atom rows (the file pipeline's default since round 2b:
// io/pipeline.py ships the TEXT; MDP_DEVICE_PARSE=0 / FrameBatches(device_parse=False) select the host parser).
//
// Why: the text of a dump is no larger than the SoA doubles parsed from it (about 30 B per "id type x y z" row in %g
// against 40 B), so shipping the TEXT over PCIe costs no more than shipping the parsed columns, and the conversion --
// the bottleneck of every file-based call once the kernels run on a B200 (host parser: ~1 ms per 100 000-atom frame
// on 16 threads, pair engine: 0.16 ms) -- disappears into a kernel that runs at memory speed.
//
// How: rows are independent (a row goes to out[slot][id - 1]; nothing needs the row's index), so no line index or prefix
// scan is built: a thread owns DP_CHUNK bytes of a frame's row section and parses every row that STARTS in its chunk,
// reading past the chunk's end as far as the row goes.  Numbers take the exact Clinger fast path of dump_line.h -- the
// very code the host parser runs, same bits.  Whatever the fast path cannot do is not approximated: the frame is
// flagged and the caller re-parses it with the host parser.  A frame is accepted only if exactly natoms rows were
// parsed, every id was in 1..natoms and no id occurred twice (bitmap with atomicOr) -- then every output element has been
// written exactly once, and it is what the host parser would have written.
#include <algorithm>

#include "common.cuh"
#include "dump_rows.h"

namespace {

constexpr int DP_THREADS = 128;

struct CudaAtomics {
    __device__ __forceinline__ unsigned or32(unsigned *p, unsigned v) const { return atomicOr(p, v); }
    __device__ __forceinline__ void add64(unsigned long long *p, unsigned long long v) const { atomicAdd(p, v); }
    __device__ __forceinline__ void or64(unsigned long long *p, unsigned long long v) const { atomicOr(p, v); }
};

__global__ void __launch_bounds__(DP_THREADS) k_dump_rows(const DevParseParams p)
{
    mdp_parse_chunk(p, (int)blockIdx.y, (long long)blockIdx.x * DP_THREADS + threadIdx.x, CudaAtomics());
}

} // namespace

extern "C" {

int mdp_dump_parse_device(mdp_ctx *ctx, int nframes, const char *text, const int64_t *begin, const int64_t *end,
                          int64_t longest, int64_t natoms, int ncols, const int *colsel, int id_col, int nwant, double *out,
                          int64_t frame_stride, int64_t out_stride, uint32_t *seen, uint64_t *status, void *stream)
{
    MDP_REQUIRE(ctx && text && begin && end && colsel && out && seen && status, "mdp_dump_parse_device: NULL argument");
    MDP_REQUIRE(nframes > 0 && nframes <= 65535 && natoms > 0 && ncols > 0 && longest >= 0, "mdp_dump_parse_device: bad sizes");
    MDP_REQUIRE(nwant > 0 && nwant <= DP_MAX_WANT, "mdp_dump_parse_device: 1..%d wanted columns", DP_MAX_WANT);
    MDP_REQUIRE(id_col >= 0 && id_col < ncols, "mdp_dump_parse_device: the dump has no id column (rows cannot be placed)");
    MDP_REQUIRE(out_stride >= natoms && frame_stride >= (int64_t)nwant * out_stride, "mdp_dump_parse_device: bad strides");
    cudaStream_t st = (cudaStream_t)stream;
    MDP_CUDA(cudaSetDevice(ctx->device));
    DevParseParams p;
    memset(&p, 0, sizeof(p));
    int last = id_col, found = 0;
    for (int c = 0; c < ncols; ++c)
        if (colsel[c] >= 0) {
            MDP_REQUIRE(colsel[c] < nwant, "mdp_dump_parse_device: column slot %d out of range", colsel[c]);
            last = c > last ? c : last;
            ++found;
        }
    MDP_REQUIRE(found == nwant, "mdp_dump_parse_device: %d of %d wanted columns are mapped", found, nwant);
    MDP_REQUIRE(last < DP_MAX_COLS, "mdp_dump_parse_device: the last needed column (%d) is beyond the first %d", last, DP_MAX_COLS);
    p.nlook = last + 1;
    for (int c = 0; c < DP_MAX_COLS; ++c) p.colsel[c] = (signed char)(c < ncols && c < p.nlook ? colsel[c] : -1);
    // no scratch of the context is used (this entry point is called from the pipeline's reader thread on its copy stream,
    // next to kernels of the same context on the consumer's stream): the caller owns seen and status
    const int64_t seen_words = (natoms + 31) / 32;
    MDP_CUDA(cudaMemsetAsync(seen, 0, (size_t)nframes * seen_words * 4, st));
    MDP_CUDA(cudaMemsetAsync(status, 0, (size_t)nframes * 16, st));
    if (longest == 0) return 0;
    p.text = text;
    p.begin = (const long long *)begin;
    p.end = (const long long *)end;
    p.natoms = natoms;
    p.id_col = id_col;
    p.nwant = nwant;
    p.out = out;
    p.frame_stride = frame_stride;
    p.out_stride = out_stride;
    p.seen = seen;
    p.seen_words = seen_words;
    p.status = (unsigned long long *)status;
    const int64_t chunks = ceil_div<int64_t>(longest, DP_CHUNK);
    dim3 grid((unsigned)ceil_div<int64_t>(chunks, DP_THREADS), (unsigned)nframes);
    k_dump_rows<<<grid, DP_THREADS, 0, st>>>(p);
    return mdp_check_launch("k_dump_rows");
}

} // extern "C"
