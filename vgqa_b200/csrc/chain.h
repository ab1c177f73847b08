// Program of a fused row-tile GEMM chain (chain.cu).  Internal header.
#pragma once
#include <vector>

#include "common.h"

namespace vg {

static constexpr int CH_MAX_OPS = 8;
static constexpr int CH_MAX_TM = 3;

enum ChOpKind { CH_GEMM = 0, CH_FFN = 1, CH_LOAD = 2 };
enum ChAKind { CH_A_ACT = 0, CH_A_STREAM = 1 };
enum ChWait { CH_WAIT_NONE = 0, CH_WAIT_TMA = 1, CH_WAIT_EPI = 2 };

// One step of a chain.  Activations live in eight 16 KB shared-memory tiles ("ACT blocks": [128 rows x 64 bf16], 128B-swizzled,
// ACT0 = blocks 0..3, ACT1 = blocks 4..7); a 256-wide activation occupies four consecutive blocks.
struct ChOp {
  int kind = CH_GEMM;
  // ---- A operand.  CH_A_ACT: `nkb` consecutive ACT blocks from a_blk; CH_A_STREAM: [128 x 64] tiles of the row-major global
  // matrix behind tensor map a_tm, columns a_col0 + 64 kb (the K = 2048 inputs).  CH_LOAD: tensor map → nkb ACT blocks from a_blk.
  int a_kind = CH_A_ACT, a_blk = 0, a_tm = 0, a_col0 = 0;
  int a_wait = CH_WAIT_NONE;   // what makes the ACT buffer of a_blk valid for THIS op (one waiter per TMA load / epilogue write)
  // ---- weights: tiles [nblk][nkb] of 128 N-rows x 64 K (chain_tile_weights); an accumulator chunk covers cn n-blocks (1 or 2)
  const bf16* w = nullptr;
  int nkb = 4, nblk = 2, cn = 2;
  int acc0 = 0;                // accumulator (0 / 1) of the first chunk; chunks alternate
  // ---- epilogue:  v = act(acc + bias[col] + table[t(row)][col]);  with ln_w:  y = LayerNorm(v + res32[row]) → out32 (fp32 [M,256])
  const float* bias = nullptr;
  const float* table = nullptr;
  int table_ld = 0;
  int act = ACT_NONE;
  const float* res32 = nullptr;
  const float* ln_w = nullptr;
  const float* ln_b = nullptr;
  float ln_eps = 1e-5f;
  float* out32 = nullptr;
  const float* ln2_w = nullptr;   // optional second LayerNorm of the normalised row → out2 (bf16 [M, ld_out2])
  const float* ln2_b = nullptr;
  float ln2_eps = 1e-5f;
  bf16* out2 = nullptr;
  int ld_out2 = 256;
  int out_act = -1;               // first ACT block that receives the bf16 rows (the next GEMM's A operand), or -1
  bf16* out_bf16 = nullptr;       // row-major global bf16 copy [M, ld_out]
  int ld_out = 0;
  float* out_f32 = nullptr;       // row-major global fp32 copy [M, ld_out_f32] (plain epilogue only)
  int ld_out_f32 = 0;
  const float* head_w = nullptr;  // plain epilogue over a 256-wide chunk: y[j] = head_act(sum_c bf16(v[c]) head_w[j][c] + head_b[j]), j < head_n <= 4
  const float* head_b = nullptr;
  int head_n = 0, head_act = 0;   // head_act 1 = sigmoid
  float* head_out = nullptr;      // [M, head_n] fp32
  // with head_n == 4 (boxes cx, cy, w, h): gen_sineembed_for_position of the head's output (model_utils.py:15-40; order y, x, w, h,
  // 128 features each) written as the 512-wide bf16 A operand of the next GEMM into ACT blocks 0..7 (and to sine_out [M, 512])
  int gen_sine = 0;
  bf16* sine_out = nullptr;
  // ---- CH_FFN:  y = LN(res32 + W2 relu(W1 x + bias) + bias2), x = ACT0; w = W1 tiles [ff_chunks][4], w2 = W2 tiles [2][2 ff_chunks]
  const bf16* w2 = nullptr;
  const float* bias2 = nullptr;
  int ff_chunks = 0;              // FFN_DIM / 128 (even)
};

struct ChainParams {
  CUtensorMap tm[CH_MAX_TM];
  int n_tm = 0;
  int n_ops = 0;
  int M = 0;    // rows (clips x frames)
  int T = 1;    // frames per clip: table row of a query row = row % T
  float sine_inv[64];   // 2*pi / 10000^(2k/128): filled by chain_launch
  ChOp ops[CH_MAX_OPS];
};

void chain_set_tmap(ChainParams& p, int i, const bf16* ptr, int rows, int cols, int ld);
void chain_launch(ChainParams& p, cudaStream_t stream);
size_t chain_tile_weights(const float* W, int N, int K, std::vector<bf16>& out);

}  // namespace vg
