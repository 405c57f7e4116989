// Shared-prefix decode attention for beam search (4 beams x 4:1 GQA).
//
// The reference expands the KV cache to one copy per beam (patch_hf.py:336-339) and every beam row re-reads all
// of it each step.  Here the beams of a sentence share the prompt's KV pages, so the prefix is read ONCE per
// (sentence, kv head): the 4 beams x 4 query heads are exactly the 16 rows of mma.sync m16n8k16 (the per-row kernel
// in decode_attention.cuh leaves 12 of those rows as padding).  The private tails (<= 4 pages = 64 keys per beam) follow
// packed into shared tiles: 16 keys of each beam side by side, each beam's 4 rows unmasked on its own 16 columns only
// (one extra tile while the tails are <= 16 keys); one launch replaces k per-row attentions, and with one key split the
// kernel writes the attention output itself (no combine pass).
// Same loader (4 warps, cp.async ring, mbarriers), same rotate-at-write key convention (q_sys variant for the pinned
// prefix tiles, ring variant after it).  HBM-bound: 4096 * prefix_len bytes per layer per SENTENCE.
#pragma once
#include "decode_attention.cuh"

namespace isst {

struct DecodeGroupParams {
  const bf16* qkv;          // [n rows, (H + 2 Hkv) * HD]: q rotated in place (ring variant) by llm_rope_append_kernel
  const bf16* q_sys;        // [n rows, H * HD]
  PagedKV kv;               // tables of the group's FIRST row are used (the shared pages are the same in all rows)
  const int* slots;         // [n rows]
  const int* key_hi;        // [n groups] logical length of the shared prefix (whole pages)
  const int* tail_page;     // [n groups] page-table index where every beam's private pages start
  bf16* out;                // [n rows][H * HD]: written directly when splits == 1
  float* part_o;            // [n rows][H][splits][HD]
  float* part_ml;           // [n rows][H][splits][2]
  int H;
  int splits;               // key splits; the last one also takes the private tails
  float scale_log2;
  // ---- fused RoPE + KV append (decode step of the chain path): the kernel completes the q rows of its 16 (beam, head)
  //      pairs from the QKV GEMM's fp32 split partials and rotates them (both variants); the CTA that owns the tails also
  //      completes, rotates and appends the new token's K / V of its kv head for the 4 beams before it reads the tails.
  //      llm_rope_append_kernel is then not launched.  fuse == 0: q / q_sys / cache were prepared by that kernel.
  int fuse;
  const float* part;        // [n_part][n rows][(H + 2 Hkv) * HD] fp32
  int n_part;
  long long part_stride;
  const float2* tab_ring;   // [n rows][HD / 2] (cos, sin) at the absolute position
  const float2* tab_sys;    // [n rows][HD / 2] at the prefix-convention position
  const int* active;        // [n rows] or null: rows that append
};

constexpr int kGrpRows = 16;                       // 4 beams x 4 query heads of one kv head
constexpr int kGrpQLds = 128 + 8;
constexpr int kGrpSmemBytes = kDecStages * kDecStageElems * 2 + 2 * kGrpRows * kGrpQLds * 2 + 2 * kDecStages * 8 + 8;

__global__ void __launch_bounds__(kDecThreads, 2)
decode_attention_group_kernel(const DecodeGroupParams p) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int HD = 128, LDS = kDecLds, TILE = kDecTile, GROUP = 4, BEAMS = 4, ROWS = kGrpRows, QL = kGrpQLds;
  extern __shared__ __align__(16) uint8_t dec_smem[];
  bf16* stage_base = reinterpret_cast<bf16*>(dec_smem);
  bf16* qbuf = stage_base + kDecStages * kDecStageElems;       // [2 variants][16 rows][QL]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(qbuf + 2 * ROWS * QL);
  uint64_t* empty_bar = full_bar + kDecStages;
  uint64_t* app_bar = empty_bar + kDecStages;                  // fused mode: the new token's K / V are in the pool

  const int split = blockIdx.x, head = blockIdx.y, grp = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int b0 = grp * BEAMS;
  const int slot = p.slots[b0];
  const int L = p.key_hi[grp];
  const int sys_len = min(p.kv.sys_len[slot], L), ring_start = p.kv.ring_start[slot];
  const int* table = p.kv.page_table + static_cast<size_t>(slot) * p.kv.pages_per_stream;
  const int n_sys_tiles = (sys_len + TILE - 1) / TILE;
  const int tiles_total = L <= 0 ? 0 : n_sys_tiles + (L - sys_len + TILE - 1) / TILE;
  const int tiles_per = (tiles_total + p.splits - 1) / p.splits;
  const int t_lo = min(tiles_total, split * tiles_per), t_hi = min(tiles_total, t_lo + tiles_per);
  const bool tails = split == p.splits - 1;          // the last split also attends to the 4 private tails
  // partial slot of (row r = beam * 4 + hq): ((b0 + beam) * H + head * 4 + hq) * splits + split
  auto pslot = [&](int r) { return (static_cast<size_t>(b0 + (r >> 2)) * p.H + head * GROUP + (r & 3)) * p.splits + split; };
  if (t_lo >= t_hi && !tails) {   // empty split: neutral partials
    for (int idx = tid; idx < ROWS * HD; idx += kDecThreads) {
      const int r = idx / HD, d = idx % HD;
      p.part_o[pslot(r) * HD + d] = 0.f;
      if (d == 0) { p.part_ml[pslot(r) * 2] = -INFINITY; p.part_ml[pslot(r) * 2 + 1] = 0.f; }
    }
    return;
  }
  const int n_pre = t_hi - t_lo;                     // shared-prefix tiles of this split
  // private keys [L, kv_len + 1) of the 4 beams, packed: tail tile u holds keys [16 u, 16 u + 16) of beam w in the
  // 16-key slice that compute warp w works on (a page per beam), so the usual <= 16 private keys cost ONE tile
  auto tail_len = [&](int beam) { return min(TILE, p.kv.kv_len[p.slots[b0 + beam]] + 1 - L); };
  int max_tail = 0;
#pragma unroll
  for (int bm = 0; bm < BEAMS; ++bm) max_tail = max(max_tail, tail_len(bm));
  const int n_tiles = n_pre + (tails ? (max_tail + 15) / 16 : 0);
  const int tail_slot0 = p.tail_page[grp] * kPageTokens;
  auto tile_j0 = [&](int t) { return t < n_sys_tiles ? t * TILE : sys_len + (t - n_sys_tiles) * TILE; };
  auto tile_j1 = [&](int t) { return t < n_sys_tiles ? min(sys_len, t * TILE + TILE) : min(L, sys_len + (t - n_sys_tiles + 1) * TILE); };

  if (tid == 0) {
    for (int s0 = 0; s0 < kDecStages; ++s0) { dec_mbar_init(&full_bar[s0], 32 * kDecLoaders); dec_mbar_init(&empty_bar[s0], 4); }
    dec_mbar_init(app_bar, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= 4) {
    // ---- loader warps: identical to decode_attention_mma_kernel ----
    constexpr int kPass = 4 / kDecLoaders;
    const int pass0 = (warp - 4) * kPass;
    const int g2 = lane >> 4, chunk = lane & 15;
    const size_t page_elems = static_cast<size_t>(2) * p.kv.kv_heads * kPageTokens * HD;
    const bf16* head_base = p.kv.pool + static_cast<size_t>(head) * kPageTokens * HD + chunk * 8;
    const size_t v_off = static_cast<size_t>(p.kv.kv_heads) * kPageTokens * HD;
    int nx_s0[kPass], nx_pa[kPass], nx_pb[kPass];
    int nx_ok[kPass];
    auto lookup = [&](int ti) {                       // ti: tile index within this CTA's list
#pragma unroll
      for (int ps = 0; ps < kPass; ++ps) {
        const int gi = 2 * (pass0 + ps) + g2;
        if (ti < n_pre) {
          const int t = t_lo + ti;
          const int jg = min(tile_j0(t) + 8 * gi, L - 1);
          nx_s0[ps] = kv_slot(jg, sys_len, ring_start);
          const int last = kv_slot(min(jg + 7, tile_j1(t) - 1 > jg ? tile_j1(t) - 1 : jg), sys_len, ring_start);
          nx_pa[ps] = table[nx_s0[ps] >> 4];
          nx_pb[ps] = table[last >> 4];
          nx_ok[ps] = tile_j1(t) - (tile_j0(t) + 8 * gi);
        } else {
          // packed tail tile u: key group gi = keys [8 (gi & 1), +8) of round u of beam gi >> 1; key i of a beam sits
          // in slot tail_slot0 + i of that row's own pages (page aligned)
          const int beam = gi >> 1, off = 16 * (ti - n_pre) + 8 * (gi & 1);
          const int* tb = p.kv.page_table + static_cast<size_t>(p.slots[b0 + beam]) * p.kv.pages_per_stream;
          nx_ok[ps] = tail_len(beam) - off;
          nx_s0[ps] = tail_slot0 + off;
          nx_pa[ps] = nx_pb[ps] = nx_ok[ps] > 0 ? tb[nx_s0[ps] >> 4] : 0;
        }
      }
    };
    lookup(0);
    for (int ti = 0; ti < n_tiles; ++ti) {
      const int stage = ti % kDecStages;
      if (ti >= kDecStages) dec_mbar_wait(&empty_bar[stage], ((ti / kDecStages) - 1) & 1);
      if (p.fuse && ti == n_pre) dec_mbar_wait(app_bar, 0);      // first tail tile: the compute warps have appended the new token
#pragma unroll
      for (int ps = 0; ps < kPass; ++ps) {
        const int gi = 2 * (pass0 + ps) + g2;
        bf16* sK = stage_base + stage * kDecStageElems + (8 * gi) * LDS + chunk * 8;
        bf16* sV = sK + TILE * LDS;
        const int n_ok = nx_ok[ps];
        const int s0 = nx_s0[ps];
        const bf16* base_a = head_base + static_cast<size_t>(nx_pa[ps]) * page_elems;
        const bf16* base_b = head_base + static_cast<size_t>(nx_pb[ps]) * page_elems;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int sl = s0 + it;
          const bf16* src = (((sl ^ s0) & ~15) == 0 ? base_a : base_b) + (sl & 15) * HD;
          const bool ok = it < n_ok;
          if (!ok) src = p.kv.pool;
          cp_async16(sK + it * LDS, src, ok ? 16 : 0);
          cp_async16(sV + it * LDS, src + v_off, ok ? 16 : 0);
        }
      }
      cp_async_arrive_noinc(&full_bar[stage]);
      if (ti + 1 < n_tiles) lookup(ti + 1);
    }
    cp_async_wait_all();
    return;
  }

  // ---- compute warps: stage the 2 x 16 query rows (both RoPE variants) ----
  if (p.fuse) {
    const int ldq = (p.H + 2 * p.kv.kv_heads) * HD;
    // split partials of elements (col + d .. d + 3) and (col + d + 64 .. + 67) of one row, summed in split order and
    // rounded like the bf16 projection output; all (<= 16) 16-byte loads of a call are in flight together
    auto quad_sum = [&](int row, int col, int d, float* a, float* bb) {
      const float* p0 = p.part + static_cast<size_t>(row) * ldq + col + d;
      float4 va[8], vb[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        va[q] = make_float4(0.f, 0.f, 0.f, 0.f); vb[q] = va[q];
        if (q < p.n_part) {
          va[q] = __ldcg(reinterpret_cast<const float4*>(p0 + q * p.part_stride));
          vb[q] = __ldcg(reinterpret_cast<const float4*>(p0 + q * p.part_stride + 64));
        }
      }
      float sa[4] = {((va[0].x + va[1].x) + va[2].x) + va[3].x, ((va[0].y + va[1].y) + va[2].y) + va[3].y,
                     ((va[0].z + va[1].z) + va[2].z) + va[3].z, ((va[0].w + va[1].w) + va[2].w) + va[3].w};
      float sb[4] = {((vb[0].x + vb[1].x) + vb[2].x) + vb[3].x, ((vb[0].y + vb[1].y) + vb[2].y) + vb[3].y,
                     ((vb[0].z + vb[1].z) + vb[2].z) + vb[3].z, ((vb[0].w + vb[1].w) + vb[2].w) + vb[3].w};
#pragma unroll
      for (int q = 4; q < 8; ++q)
        if (q < p.n_part) {
          sa[0] += va[q].x; sa[1] += va[q].y; sa[2] += va[q].z; sa[3] += va[q].w;
          sb[0] += vb[q].x; sb[1] += vb[q].y; sb[2] += vb[q].z; sb[3] += vb[q].w;
        }
      for (int q = 8; q < p.n_part; ++q) {
        const float4 xa = __ldcg(reinterpret_cast<const float4*>(p0 + q * p.part_stride));
        const float4 xb = __ldcg(reinterpret_cast<const float4*>(p0 + q * p.part_stride + 64));
        sa[0] += xa.x; sa[1] += xa.y; sa[2] += xa.z; sa[3] += xa.w;
        sb[0] += xb.x; sb[1] += xb.y; sb[2] += xb.z; sb[3] += xb.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { a[j] = bf16_round(sa[j]); bb[j] = bf16_round(sb[j]); }
    };
    auto store4 = [](bf16* dst, float x0, float x1, float x2, float x3) {
      *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(x0, x1), pack_bf16(x2, x3));
    };
    // q: 16 rows x 16 groups of 4 (d, d + 64) pairs over the 128 compute threads
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
      const int item = it * 128 + tid;
      const int r = item >> 4, d = (item & 15) * 4;
      const int row = b0 + (r >> 2);
      float a[4], bb[4];
      const float4* tr = reinterpret_cast<const float4*>(p.tab_ring + static_cast<size_t>(row) * 64 + d);
      const float4* ts = reinterpret_cast<const float4*>(p.tab_sys + static_cast<size_t>(row) * 64 + d);
      const float4 r01 = tr[0], r23 = tr[1], s01 = ts[0], s23 = ts[1];      // (cos, sin) of d, d + 1 | d + 2, d + 3
      quad_sum(row, (head * GROUP + (r & 3)) * HD, d, a, bb);
      const float crc[4] = {r01.x, r01.z, r23.x, r23.z}, crs[4] = {r01.y, r01.w, r23.y, r23.w};
      const float csc[4] = {s01.x, s01.z, s23.x, s23.z}, css[4] = {s01.y, s01.w, s23.y, s23.w};
      store4(qbuf + r * QL + d, a[0] * crc[0] - bb[0] * crs[0], a[1] * crc[1] - bb[1] * crs[1], a[2] * crc[2] - bb[2] * crs[2], a[3] * crc[3] - bb[3] * crs[3]);
      store4(qbuf + r * QL + d + 64, bb[0] * crc[0] + a[0] * crs[0], bb[1] * crc[1] + a[1] * crs[1], bb[2] * crc[2] + a[2] * crs[2], bb[3] * crc[3] + a[3] * crs[3]);
      store4(qbuf + (ROWS + r) * QL + d, a[0] * csc[0] - bb[0] * css[0], a[1] * csc[1] - bb[1] * css[1], a[2] * csc[2] - bb[2] * css[2], a[3] * csc[3] - bb[3] * css[3]);
      store4(qbuf + (ROWS + r) * QL + d + 64, bb[0] * csc[0] + a[0] * css[0], bb[1] * csc[1] + a[1] * css[1], bb[2] * csc[2] + a[2] * css[2], bb[3] * csc[3] + a[3] * css[3]);
    }
    if (tails) {
      // the new token's K (rotated) and V of this kv head, for the 4 beams: 2 x 4 x 16 groups = one item per thread
      const int is_v = tid >> 6, beam = (tid >> 4) & 3, d = (tid & 15) * 4;
      const int row = b0 + beam;
      float a[4], bb[4];
      quad_sum(row, (p.H + is_v * p.kv.kv_heads + head) * HD, d, a, bb);
      if (!p.active || p.active[row]) {
        const int sl = p.slots[row];
        const int L_old = p.kv.kv_len[sl];
        const int* tb = p.kv.page_table + static_cast<size_t>(sl) * p.kv.pages_per_stream;
        const size_t off = kv_offset(p.kv, tb, kv_slot(L_old, p.kv.sys_len[sl], p.kv.ring_start[sl]), is_v, head);
        if (is_v) {
          store4(p.kv.pool + off + d, a[0], a[1], a[2], a[3]);
          store4(p.kv.pool + off + d + 64, bb[0], bb[1], bb[2], bb[3]);
        } else {
          const float4* tc = reinterpret_cast<const float4*>((L_old < p.kv.sys_len[sl] ? p.tab_sys : p.tab_ring) + static_cast<size_t>(row) * 64 + d);
          const float4 c01 = tc[0], c23 = tc[1];
          const float cc[4] = {c01.x, c01.z, c23.x, c23.z}, sn[4] = {c01.y, c01.w, c23.y, c23.w};
          store4(p.kv.pool + off + d, a[0] * cc[0] - bb[0] * sn[0], a[1] * cc[1] - bb[1] * sn[1], a[2] * cc[2] - bb[2] * sn[2], a[3] * cc[3] - bb[3] * sn[3]);
          store4(p.kv.pool + off + d + 64, bb[0] * cc[0] + a[0] * sn[0], bb[1] * cc[1] + a[1] * sn[1], bb[2] * cc[2] + a[2] * sn[2], bb[3] * cc[3] + a[3] * sn[3]);
        }
      }
      dec_mbar_arrive(app_bar);
    }
  } else {
    const int ldq = (p.H + 2 * p.kv.kv_heads) * HD;
    // 2 variants x 16 rows x 16 chunks of 16 bytes = 512 chunks over 128 threads
    for (int c = tid; c < 2 * ROWS * 16; c += 128) {
      const int v = c >> 8, r = (c >> 4) & 15, ch = c & 15;
      const int row = b0 + (r >> 2), hq = r & 3;
      const bf16* src = v == 0 ? p.qkv + static_cast<size_t>(row) * ldq + (head * GROUP + hq) * HD
                               : p.q_sys + static_cast<size_t>(row) * (p.H * HD) + (head * GROUP + hq) * HD;
      *reinterpret_cast<uint4*>(qbuf + (v * ROWS + r) * QL + ch * 8) = *reinterpret_cast<const uint4*>(src + ch * 8);
    }
  }
  dec_bar_compute();

  float o0[8][4], o1[8][4];      // O^T for queries 0-7 / 8-15: (dim g, q 2*t4) (dim g, q 2*t4+1) (dim g+8, ..) (dim g+8, ..)
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) {
    o0[mt][0] = o0[mt][1] = o0[mt][2] = o0[mt][3] = 0.f;
    o1[mt][0] = o1[mt][1] = o1[mt][2] = o1[mt][3] = 0.f;
  }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};      // rows g and g + 8 of S

  for (int ti = 0; ti < n_tiles; ++ti) {
    const int t = t_lo + ti;
    const bool is_tail = ti >= n_pre;                // packed tail tile: this warp's 16 keys belong to beam `warp`
    dec_mbar_wait(&full_bar[ti % kDecStages], (ti / kDecStages) & 1);
    const bf16* sK = stage_base + (ti % kDecStages) * kDecStageElems;
    const bf16* sV = sK + TILE * LDS;
    const bf16* qv = qbuf + ((!is_tail && t < n_sys_tiles) ? ROWS * QL : 0);
    // ---- S = Q K^T: 16 query rows x this warp's 16 keys ----
    float s[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
    {
      const bf16* krow = sK + (16 * warp + ((lane >> 4) << 3) + (lane & 7)) * LDS + (((lane >> 3) & 1) << 3);
      // A fragment by ldmatrix.x4: (rows 0-7, k lo) (rows 8-15, k lo) (rows 0-7, k hi) (rows 8-15, k hi)
      const bf16* qrow = qv + ((lane & 7) + (((lane >> 3) & 1) << 3)) * QL + ((lane >> 4) << 3);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        uint32_t kf[4], qa[4];
        ldmatrix_x4(kf, krow + kk * 16);
        ldmatrix_x4(qa, qrow + kk * 16);
        mma_bf16_16816(s[0], qa, kf[0], kf[1]);
        mma_bf16_16816(s[1], qa, kf[2], kf[3]);
      }
    }
    // ---- online softmax on rows g (elements 0, 1) and g + 8 (elements 2, 3) ----
    const int jw = is_tail ? 16 * (ti - n_pre) : tile_j0(t) + 16 * warp;
    const int j1 = is_tail ? tail_len(warp) : tile_j1(t);
    float corr[2], pr[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float mx = m_run[h];
      const bool row_on = !is_tail || ((g + 8 * h) >> 2) == warp;
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = jw + n * 8 + 2 * t4 + e;
          s[n][2 * h + e] = (row_on && j < j1) ? s[n][2 * h + e] * p.scale_log2 : -INFINITY;
          mx = fmaxf(mx, s[n][2 * h + e]);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float msafe = (mx == -INFINITY) ? 0.f : mx;
      corr[h] = (m_run[h] == -INFINITY) ? 0.f : exp2_fast(m_run[h] - msafe);
      m_run[h] = mx;
      pr[h][0] = exp2_fast(s[0][2 * h] - msafe); pr[h][1] = exp2_fast(s[0][2 * h + 1] - msafe);
      pr[h][2] = exp2_fast(s[1][2 * h] - msafe); pr[h][3] = exp2_fast(s[1][2 * h + 1] - msafe);
      l_run[h] = l_run[h] * corr[h] + ((pr[h][0] + pr[h][1]) + (pr[h][2] + pr[h][3]));
    }
    // P^T B fragments: keys (2*t4, 2*t4+1) and (8 + 2*t4, +1) of query g (tile 0) / g + 8 (tile 1)
    const uint32_t pb00 = pack_bf16(pr[0][0], pr[0][1]), pb01 = pack_bf16(pr[0][2], pr[0][3]);
    const uint32_t pb10 = pack_bf16(pr[1][0], pr[1][1]), pb11 = pack_bf16(pr[1][2], pr[1][3]);
    // O^T columns are queries 2*t4, 2*t4+1 of each tile: their corrections live in the lanes that own those rows
    const float c0_lo = __shfl_sync(0xffffffffu, corr[0], (2 * t4) * 4), c0_hi = __shfl_sync(0xffffffffu, corr[0], (2 * t4 + 1) * 4);
    const float c1_lo = __shfl_sync(0xffffffffu, corr[1], (2 * t4) * 4), c1_hi = __shfl_sync(0xffffffffu, corr[1], (2 * t4 + 1) * 4);
    const bool rescale = __any_sync(0xffffffffu, corr[0] != 1.f || corr[1] != 1.f);
    const bf16* vrow = sV + (16 * warp + ((lane >> 4) << 3) + (lane & 7)) * LDS + (((lane >> 3) & 1) << 3);
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      uint32_t va[4];
      ldmatrix_x4_trans(va, vrow + mt * 16);
      if (rescale) {
        o0[mt][0] *= c0_lo; o0[mt][1] *= c0_hi; o0[mt][2] *= c0_lo; o0[mt][3] *= c0_hi;
        o1[mt][0] *= c1_lo; o1[mt][1] *= c1_hi; o1[mt][2] *= c1_lo; o1[mt][3] *= c1_hi;
      }
      mma_bf16_16816(o0[mt], va, pb00, pb01);
      mma_bf16_16816(o1[mt], va, pb10, pb11);
    }
    __syncwarp();
    if (lane == 0) dec_mbar_arrive(&empty_bar[ti % kDecStages]);
  }
  dec_bar_compute();             // every tile has landed and been consumed: the stage buffers become the merge area

  // ---- merge the 4 warps (each saw a quarter of every tile) ----
  float* sm_o = reinterpret_cast<float*>(dec_smem);            // [4 warps][16 rows][HD]
  float* sm_m = sm_o + 4 * ROWS * HD;                          // [4][16]
  float* sm_l = sm_m + 4 * ROWS;
  float* sm_w = sm_l + 4 * ROWS;                               // [4][16] merge weights, then [16] 1 / l
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
    if (t4 == 0) { sm_m[warp * ROWS + g + 8 * h] = m_run[h]; sm_l[warp * ROWS + g + 8 * h] = l_run[h]; }
  }
#pragma unroll
  for (int mt = 0; mt < 8; ++mt) {
    float* w0 = sm_o + (warp * ROWS + 2 * t4) * HD + mt * 16 + g;
    w0[0] = o0[mt][0]; w0[HD] = o0[mt][1]; w0[8] = o0[mt][2]; w0[HD + 8] = o0[mt][3];
    float* w1 = sm_o + (warp * ROWS + 8 + 2 * t4) * HD + mt * 16 + g;
    w1[0] = o1[mt][0]; w1[HD] = o1[mt][1]; w1[8] = o1[mt][2]; w1[HD + 8] = o1[mt][3];
  }
  dec_bar_compute();
  if (tid < ROWS) {
    const int r = tid;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) mm = fmaxf(mm, sm_m[w * ROWS + r]);
    const float ms = (mm == -INFINITY) ? 0.f : mm;
    float ll = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float wm = sm_m[w * ROWS + r];
      const float wgt = (wm == -INFINITY) ? 0.f : exp2_fast(wm - ms);
      sm_w[w * ROWS + r] = wgt;
      ll += wgt * sm_l[w * ROWS + r];
    }
    if (p.splits > 1) { p.part_ml[pslot(r) * 2] = mm; p.part_ml[pslot(r) * 2 + 1] = ll; }
    sm_w[4 * ROWS + r] = ll > 0.f ? 1.f / ll : 0.f;
  }
  dec_bar_compute();
#pragma unroll 4
  for (int r = 0; r < ROWS; ++r) {
    const int d = tid;                                         // 128 compute threads == HD
    float oo = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) oo += sm_w[w * ROWS + r] * sm_o[(w * ROWS + r) * HD + d];
    if (p.splits == 1) p.out[(static_cast<size_t>(b0 + (r >> 2)) * p.H + head * GROUP + (r & 3)) * HD + d] = __float2bfloat16_rn(oo * sm_w[4 * ROWS + r]);
    else p.part_o[pslot(r) * HD + d] = oo;
  }
}

}  // namespace isst
