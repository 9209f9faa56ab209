// Sorted-index segmented reduction + sparse row update (shared by every
// training path).  Replaces grad_sum_matrix + Sm.dot(G)/n (skge/util.py:53-101)
// and ParameterUpdate.__call__ (skge/param.py:108-174).
#pragma once
#include "common.cuh"

namespace skge {

static constexpr int kMaxRoles = 6;

// How the occurrences of rows in a minibatch map to per-unit gradient rows.
// Unit i (a pair or an example) owns `rows_per_unit` rows in G; role r says
// "row idx[r][i] of table is_rel[r] receives gsign[r] * G[i][grow[r]]".
// twin[r] = q >= 0 pairs two roles whose ids often coincide within a unit (the subject of a
// positive triple and of its object-corrupted negative): when idx[r][i] == idx[q][i] the producer
// has already summed both contributions into the row of the LOWER role, which then counts as two
// occurrences in the mean, and the higher role's occurrence is dropped.
struct RoleMap {
  const int32_t *idx[kMaxRoles];
  int is_rel[kMaxRoles];
  int grow[kMaxRoles];
  float gsign[kMaxRoles];
  int twin[kMaxRoles] = {-1, -1, -1, -1, -1, -1};
  int nroles;
  // Run weights (nullable): the producer has pre-summed the rows that role `runw_role` of several units
  // would send to the same table row (consecutive pairs of one relation) into the row of the run's first
  // unit.  runw[i] > 0: unit i carries such a sum standing for runw[i] occurrences; runw[i] == 0: unit i's
  // occurrence of that role is contained in an earlier unit's row and is dropped.
  const int32_t *runw = nullptr;
  int runw_role = -1;
  // Shared rows (nullable; HolE pairs whose negative keeps the relation and corrupts exactly one
  // entity): two of the pair's three entity gradients are the same row times gp resp. gn, so the
  // producer writes that row once, unscaled, and (gp, gn) into coef[2 i], coef[2 i + 1].  With roles
  // ordered (sp, sn, op, on, ...): object corrupted (sp == sn): roles 2, 3 read row 2 times gp, gn
  // (payload codes 6, 7); subject corrupted (op == on): roles 0, 1 read row 0 times gp, gn (codes 14, 15).
  const float *coef = nullptr;
};

// order[] = the indices 0..n-1 sorted (stably) by keys[] (values below 1 << key_bits)
size_t order_workspace_bytes(int64_t n);
int order_by_key(const int32_t *keys, int64_t n, int key_bits, Arena &ar, cudaStream_t st, const int32_t **order);

// One parameter table (E, or the relation table R).
struct ParamDesc {
  float *param;        // table, row-major [rows][d]
  float *p2;           // AdaGrad accumulator (UPDATE + ADAGRAD only)
  int post;            // SKGE_POST_*
  float rparam;        // g += rparam * row  (skge/hole.py:33,40,83)
  int32_t *upd_counts; // nullable
  float *out_g;        // emit mode: [U][d]
  int32_t *out_idx;    // emit mode: [U]
  float *hat;          // spectral mode: packed spectra of the table's rows, refreshed on update (nullable)
};

size_t seg_workspace_bytes(int64_t L, int d);

// Builds keys from `rm` (masked by flags, nullable), sorts, finds segments and
// either emits (mean gradient, row id) per unique row (update == false) or
// applies the optimiser step in place (update == true).  counts[1], counts[2]
// receive the number of unique rows of table 0 / table 1.
// spectral != 0: the rows of G are packed spectra (csrc/fft.cuh); each unique row's summed
// spectrum is taken back to the time domain before the update and the updated row's spectrum
// is written to ParamDesc::hat (HolE in the frequency domain, update mode, power-of-two d).
int seg_run(const RoleMap &rm, const uint8_t *flags, int64_t P, int64_t N, int64_t M, int d,
            const float *G, int rows_per_unit, const ParamDesc pd[2], bool update, int opt, float lr,
            int32_t *counts, Arena &ar, cudaStream_t st, int spectral = 0);

// *p = v on the stream (keeps the step capturable in a CUDA graph)
int set_i32(int32_t *p, int32_t v, cudaStream_t st);

// Sort-only service for callers with their own reduction (RESCAL's W gradient):
// sorted unit ids grouped by key, with segment starts/keys and meta[0] = nseg.
struct SegLists {
  const int32_t *vals;       // sorted payload (unit * 16 + (weight - 1) * 8 + role)
  const int32_t *seg_start;  // [nseg + 1]
  const int32_t *seg_key;    // [nseg]
  const int32_t *meta;       // [0] = nseg, [1] = segments of table 0
};
int seg_build(const RoleMap &rm, const uint8_t *flags, int64_t P, int64_t N, int64_t M, Arena &ar,
              cudaStream_t st, SegLists *out);

}  // namespace skge
