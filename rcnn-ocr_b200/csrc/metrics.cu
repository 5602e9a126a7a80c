// K5: batched edit distance on the device for the validation metrics.
//
// Replaces the per-pair host loop of training/train.py:582-598 / evaluate_dataset.py:104-119 over
// training/metrics.py:5-32: CER = Levenshtein(reference, hypothesis) / len(reference) on CHARACTERS,
// WER = the same on WORDS (jiwer's default: collapse / strip whitespace, split on spaces), accuracy =
// exact string match.  Hypotheses arrive as the class ids the greedy decoder (K4) left on the device,
// references as the CTC target ids; both are expanded to Unicode code points through the charset table
// (class k -> the code points of itos[k-1]; a multi-character token such as "<PAD>" expands to all
// its characters), so the distances equal the reference's string distances exactly.
//
// One warp per (reference, hypothesis) pair: lane 0 expands the two sequences into shared memory (for
// WER: into 64-bit FNV-1a hashes of the words), all lanes then sweep the DP matrix by anti-diagonals
// (three rolling diagonals in shared memory).  Integer work, bit-exact.
#include "common.cuh"

namespace rcnn {
namespace {

constexpr int kMaxLen = 320;        // expanded symbols per sequence (T = 64 frames x 5-character tokens)
constexpr int kWarps = 4;

struct EditParams {
    const int32_t *hyp;  long long hyp_stride;  const int32_t *hyp_len;     // [N, hyp_stride], [N]
    const long long *ref;  const long long *ref_off;  const long long *ref_len;   // flat ids, [N] offsets, [N] lengths
    const int32_t *cp_off;  const int32_t *cp;  int C;                       // class k -> cp[cp_off[k] .. cp_off[k+1])
    int N, words;
    int32_t *dist, *nref, *nhyp;                                              // [N] each; dist = -1: sequence too long
};

struct WarpScratch {
    unsigned long long a[kMaxLen], b[kMaxLen];   // reference / hypothesis symbols (code points or word hashes)
    int diag[3][kMaxLen + 1];
};

// lane 0: class ids -> symbols.  Returns the symbol count or -1 when it does not fit.
template <typename IdT>
__device__ int expand(const IdT *ids, int n, const EditParams &p, unsigned long long *out) {
    int len = 0;
    if (!p.words) {
        for (int i = 0; i < n; ++i) {
            const long long k = (long long)ids[i];
            if (k < 0 || k >= p.C) continue;             // padding / out-of-range ids carry no characters
            for (int j = p.cp_off[k]; j < p.cp_off[k + 1]; ++j) {
                if (len >= kMaxLen) return -1;
                out[len++] = (unsigned long long)(unsigned)p.cp[j];
            }
        }
        return len;
    }
    unsigned long long h = 1469598103934665603ull;      // FNV-1a over the code points of the current word
    bool in_word = false;
    for (int i = 0; i < n; ++i) {
        const long long k = (long long)ids[i];
        if (k < 0 || k >= p.C) continue;
        for (int j = p.cp_off[k]; j < p.cp_off[k + 1]; ++j) {
            const unsigned c = (unsigned)p.cp[j];
            const bool space = c == 32u;                 // jiwer's default word delimiter
            if (space) {
                if (in_word) {
                    if (len >= kMaxLen) return -1;
                    out[len++] = h;
                    h = 1469598103934665603ull;
                    in_word = false;
                }
            } else {
                for (int byte = 0; byte < 4; ++byte) { h ^= (c >> (8 * byte)) & 0xffu; h *= 1099511628211ull; }
                in_word = true;
            }
        }
    }
    if (in_word) {
        if (len >= kMaxLen) return -1;
        out[len++] = h;
    }
    return len;
}

__global__ void __launch_bounds__(32 * kWarps) edit_distance_kernel(const EditParams p) {
    extern __shared__ unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch &s = reinterpret_cast<WarpScratch *>(smem_raw)[warp];
    for (int n = blockIdx.x * kWarps + warp; n < p.N; n += gridDim.x * kWarps) {
        int la = 0, lb = 0;
        if (lane == 0) {
            la = expand(p.ref + p.ref_off[n], (int)p.ref_len[n], p, s.a);
            lb = expand(p.hyp + (long long)n * p.hyp_stride, p.hyp_len[n], p, s.b);
        }
        la = __shfl_sync(FULL, la, 0);
        lb = __shfl_sync(FULL, lb, 0);
        __syncwarp();
        if (la < 0 || lb < 0) {
            if (lane == 0) { p.dist[n] = -1; p.nref[n] = la; p.nhyp[n] = lb; }
            continue;
        }
        // D[i][j] = distance(a[:i], b[:j]); diagonal d holds D[i][d-i] at index i
        int *d2 = s.diag[0], *d1 = s.diag[1], *d0 = s.diag[2];
        for (int d = 0; d <= la + lb; ++d) {
            const int ilo = max(0, d - lb), ihi = min(la, d);
            for (int i = ilo + lane; i <= ihi; i += 32) {
                const int j = d - i;
                int v;
                if (i == 0) v = j;
                else if (j == 0) v = i;
                else v = min(min(d1[i - 1] + 1, d1[i] + 1), d2[i - 1] + (s.a[i - 1] != s.b[j - 1] ? 1 : 0));
                d0[i] = v;
            }
            __syncwarp();
            int *t = d2; d2 = d1; d1 = d0; d0 = t;
        }
        if (lane == 0) { p.dist[n] = d1[la]; p.nref[n] = la; p.nhyp[n] = lb; }
        __syncwarp();
    }
}

}  // namespace
}  // namespace rcnn

extern "C" int rcnn_edit_distance(const int32_t *hyp_ids, int64_t hyp_stride, const int32_t *hyp_len,
                                  const int64_t *ref_ids, const int64_t *ref_off, const int64_t *ref_len, int N,
                                  const int32_t *cp_off, const int32_t *cp, int C, int words,
                                  int32_t *dist_out, int32_t *nref_out, int32_t *nhyp_out, rcnn_stream_t stream) {
    using namespace rcnn;
    RCNN_CHECK_ARG(N >= 0 && C >= 1, "edit_distance: bad shape N=%d C=%d", N, C);
    if (N == 0) return RCNN_OK;
    RCNN_CHECK_ARG(hyp_ids && hyp_len && ref_ids && ref_off && ref_len && cp_off && cp && dist_out && nref_out && nhyp_out,
                   "edit_distance: null pointer");
    EditParams p;
    p.hyp = hyp_ids; p.hyp_stride = hyp_stride; p.hyp_len = hyp_len;
    p.ref = (const long long *)ref_ids; p.ref_off = (const long long *)ref_off; p.ref_len = (const long long *)ref_len;
    p.cp_off = cp_off; p.cp = cp; p.C = C;
    p.N = N; p.words = words != 0;
    p.dist = dist_out; p.nref = nref_out; p.nhyp = nhyp_out;
    const size_t smem = sizeof(WarpScratch) * kWarps;
    RCNN_CUDA(cudaFuncSetAttribute(edit_distance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int blocks = (N + kWarps - 1) / kWarps < 148 * 4 ? (N + kWarps - 1) / kWarps : 148 * 4;
    edit_distance_kernel<<<blocks, 32 * kWarps, smem, (cudaStream_t)stream>>>(p);
    RCNN_LAUNCH_CHECK("edit_distance_kernel");
    return RCNN_OK;
}
