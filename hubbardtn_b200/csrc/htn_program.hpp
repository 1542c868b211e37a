// Contraction programs: a fixed sequence of grouped-GEMM and mix stages over slot-addressed
// block-sparse tensors.  Every operator of the hot path (H_AC, H_C, H_AC2, the MPO transfers,
// blockwise products, Gram/projection steps of the gauge fixing) is lowered ONCE on the host to
// such a program -- this is where the fusion-tree bookkeeping TensorKit/TensorOperations redo on
// every contraction call is spent -- and then replayed with a handful of
// launches per application.
#pragma once
#include <string>
#include <vector>

#include "htn_internal.hpp"

namespace htn {

// host-side operand reference: slot >= 0 -> tensor bound at launch; SLOT_WS -> program workspace
constexpr int SLOT_WS = -1;
struct Opnd {
  int slot;
  int64_t off;
};

struct GemmSegH {
  Opnd A;
  int lda;
  Opnd B;
  int ldb;
  int K;
};
// C[M x N] (row-major, ldc) = sum_segs A[M x K] * B[K x N]
struct GemmTaskH {
  Opnd C;
  int ldc, M, N;
  std::vector<GemmSegH> segs;
};
struct MixSrcH {
  Opnd src;
  double coef;
};
// dst[0 .. nelem) = sum coef * src[0 .. nelem)   (flat padded ranges of identical layout)
struct MixTaskH {
  Opnd dst;
  int nelem;
  std::vector<MixSrcH> srcs;
  int wave = -1;  // fused stage only: wave of stack jobs whose T blocks this target reads
};
// one job of the stacked stage L (htn_stackl.cuh): C[M x nt] = A[M x K] . B[K x nt]
struct StackJobH {
  Opnd A;
  int lda;
  Opnd B;
  int ldb;
  Opnd C;
  int ldc;
  int K, nt, nb, M;
  int tmap, arow;  // tensor map index / first row inside that panel (tmap < 0: plain pointer path)
  int wave;        // >= 0: the whole job belongs to this wave (one arrival per consumer warp at the end of the job)
  // tile-level waves (wave < 0): tile t of the job covers T blocks of the waves tw[t].x .. tw[t].y; every consumer warp of
  // the group that computed the tile reports to each of them when the tile is stored
  std::vector<std::pair<int, int>> tw;
};

enum StageTag { TAG_L = 1, TAG_W = 2, TAG_R = 4, TAG_Y = 8 };

struct Stage {
  int kind;  // 0 gemm, 1 mix, 2 stacked gemm with the mix riding along (htn_stackl.cuh)
  int tag;
  // host tables (kept until finalize)
  std::vector<GemmItem> items;
  std::vector<GemmSeg> segs;
  std::vector<MixTarget> mt;
  std::vector<MixSrc> ms;
  std::vector<MixChunk> mc;
  // device tables
  GemmItem* d_items = nullptr;
  GemmSeg* d_segs = nullptr;
  MixTarget* d_mt = nullptr;
  MixSrc* d_ms = nullptr;
  MixChunk* d_mc = nullptr;
  MixChunkX* d_mcx = nullptr;
  int n = 0;  // items or chunks
  int grid = 0;
  // kind 2
  std::vector<StackJob> sjobs;
  std::vector<int> wave_need;
  StackJob* d_sjobs = nullptr;
  int* d_wave_need = nullptr;
  std::vector<int2> tile_waves;
  int2* d_tile_waves = nullptr;
  unsigned long long* d_ctr = nullptr;
  int n_sjobs = 0, nwaves = 0;
  int tmap_slot = -1;  // slot whose tensor supplies the tensor maps at launch (-1: none needed)
  mutable unsigned long long epoch = 0;
};

struct Program {
  htn_ctx* ctx = nullptr;
  int nslots = 0;
  std::vector<Stage> stages;
  double* ws = nullptr;
  int64_t ws_elems = 0;
  double flops = 0, padded_flops = 0;
  double flops_tag[16] = {0};
  int n_gemm_tiles_tag[16] = {0};
  bool finalized = false;

  // ---- building ----
  int64_t ws_alloc(int64_t elems);  // 16-element aligned, zero-initialised at finalize
  // direct stage: every task's C is overwritten (tasks without segments are zero-filled)
  void add_gemm(std::vector<GemmTaskH>& tasks, int tag);
  // reduce stage: long K loops are split, partial tiles go to workspace, and a following mix stage
  // sums partials (fixed order => deterministic) plus `extra[i]` into task i's C.  Tasks with
  // neither segments nor extra sources are zero-filled.
  // adaptive_split: cut the K loops finer when the stage has too few tile-parts for the persistent grid (H_eff stage R)
  void add_gemm_reduce(std::vector<GemmTaskH>& tasks, std::vector<std::vector<MixSrcH>>& extra, int tag_gemm,
                       int tag_mix, bool adaptive_split = false);
  void add_mix(std::vector<MixTaskH>& tasks, int tag);
  // fused stage: stack jobs (ordered by wave) + the mix targets that consume their outputs
  void add_stack(std::vector<StackJobH>& jobs, std::vector<MixTaskH>& mixes, int nwaves, int tmap_slot, int tag);
  int32_t finalize(htn_ctx* ctx, int nslots);
  // ---- running ----
  int32_t run(const double* const* slots, int mask = 0xF, const unsigned char* const* slot_tmaps = nullptr) const;
  int launches(int mask = 0xF) const;
  void destroy();
};

}  // namespace htn

// Generic plan handle of the C ABI: a program + the tensors bound to its slots by default.
struct htn_plan {
  htn_ctx* ctx = nullptr;
  int kind = 0;  // HTN_PLAN_*
  htn::Program prog;
  htn_tensor* like_in = nullptr;   // structural copies (no data use)
  htn_tensor* like_out = nullptr;
  const htn_tensor* bound[htn::MAX_SLOTS] = {nullptr};  // plan-lifetime operands (GL, GR, ...)
  double stats[12] = {0};
  htn_tensor* hx = nullptr;  // host-path staging
  htn_tensor* hy = nullptr;
};

#define HTN_PLAN_HEFF_AC 1
#define HTN_PLAN_HEFF_C 2
#define HTN_PLAN_TRANSFER_L 3
#define HTN_PLAN_TRANSFER_R 4
#define HTN_PLAN_HEFF_AC2 5
