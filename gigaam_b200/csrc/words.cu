// Word grouping of greedy hypotheses on the device (gigaam/timestamps_utils.py:13-53 frames_to_words; the per-utterance
// Python loop of gigaam/model.py:104-124).  The greedy kernels already leave (token id, frame) pairs in device memory;
// this kernel turns them into word records -- first / last token frame and the token range -- so the host only
// multiplies by the frame shift and joins the pieces of each range.
//
// Reference semantics, restated on a per-token flag table built once from the tokenizer:
//   bit 0  delimiter: the piece is exactly " " -> closes the open word, contributes nothing
//   bit 1  the piece starts with U+2581 -> closes the open word, then joins the new one (without the prefix)
//   bit 2  the piece (prefix removed) is empty after strip(): it still extends the word's frame span, but a word made
//          only of such pieces is dropped (its text is empty after strip())
// A word starts at the frame of its first piece and ends one frame after its last piece.
//
// One warp per utterance: 32 tokens per coalesced load, flags turned into ballots, the (warp-uniform) word state
// machine walks the set bits; lane 0 writes the records.
#include "kernels.h"

namespace gam {
namespace {

__global__ void __launch_bounds__(128) group_words_kernel(const int* __restrict__ ids, const int* __restrict__ frames,
                                                          const int* __restrict__ counts, const unsigned char* __restrict__ flags,
                                                          int B, int V, int max_out, int max_words, int* __restrict__ w_start,
                                                          int* __restrict__ w_end, int* __restrict__ w_first, int* __restrict__ w_ntok,
                                                          int* __restrict__ n_words) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int n = min(max(counts[b], 0), max_out);
  const int* idr = ids + static_cast<size_t>(b) * max_out;
  const int* frr = frames + static_cast<size_t>(b) * max_out;
  const size_t wbase = static_cast<size_t>(b) * max_words;
  bool open = false, visible = false;
  int first_tok = 0, first_frame = 0, last_frame = 0, ntok = 0, nw = 0;
  auto commit = [&]() {
    if (open && visible) {
      if (lane == 0 && nw < max_words) {
        w_start[wbase + nw] = first_frame;
        w_end[wbase + nw] = last_frame + 1;
        w_first[wbase + nw] = first_tok;
        w_ntok[wbase + nw] = ntok;
      }
      ++nw;
    }
    open = false;
  };
  for (int base = 0; base < n; base += 32) {
    const int j = base + lane;
    const bool in = j < n;
    const int tok = in ? idr[j] : 0;
    const int fr = in ? frr[j] : 0;
    const unsigned fl = (in && tok >= 0 && tok < V) ? flags[tok] : 0u;
    const unsigned valid = __ballot_sync(0xffffffffu, in);
    const unsigned delim = __ballot_sync(0xffffffffu, (fl & 1u) != 0);
    const unsigned start = __ballot_sync(0xffffffffu, (fl & 2u) != 0);
    const unsigned blank = __ballot_sync(0xffffffffu, (fl & 4u) != 0);
    for (unsigned m = valid; m != 0; m &= m - 1) {
      const int i = __ffs(m) - 1;
      const unsigned bit = 1u << i;
      if (delim & bit) { commit(); continue; }
      if (start & bit) commit();
      const int f = __shfl_sync(0xffffffffu, fr, i);
      if (!open) { open = true; visible = false; first_tok = base + i; first_frame = f; ntok = 0; }
      last_frame = f;
      ++ntok;
      visible = visible || !(blank & bit);
    }
  }
  commit();
  if (lane == 0) n_words[b] = nw;
}

}  // namespace

void launch_group_words(const int* ids, const int* frames, const int* counts, const unsigned char* flags, int B, int V, int max_out,
                        int max_words, int* w_start, int* w_end, int* w_first, int* w_ntok, int* n_words, cudaStream_t s) {
  if (B <= 0) return;
  group_words_kernel<<<(B + 3) / 4, 128, 0, s>>>(ids, frames, counts, flags, B, V, max_out, max_words, w_start, w_end, w_first, w_ntok,
                                                 n_words);
}

}  // namespace gam
