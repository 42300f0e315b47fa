// Witness solver: what gnark's spr.Solve does between frontend.NewWitness and the prover's
// first round (/root/reference/algoplonk.go:81-89) -- every PLONK row  ql*a + qr*b + qm*a*b + qo*c + qk = 0  with
// exactly one unassigned wire determines that wire -- as a LEVEL-PARALLEL pass on the GPU.
//
//  * create(): one host pass over the rows in order finds, for every row, the wire it solves (if any) and the row's
//    level = 1 + the deepest level among the wires it reads; rows are counting-sorted by level.  gnark computes the
//    same levels at compile time and runs the rows of a level on goroutines; here a level is one kernel launch
//    (one thread per row), and runs of consecutive NARROW levels share one single-block launch that steps through
//    them with __syncthreads() -- a dependency chain costs one barrier per link instead of one launch.
//  * solve(): inputs up, levels, then one gather kernel writes L, R, O (n rows, padding rows = variable 0 as
//    gnark's NewTrace pads) and checks EVERY row's gate equation; L, R, O stay in HBM for b2p_prove_dev.
//  * A chain is still a chain: a depth-2^20 squaring chain takes ~2 us per link on one SM, so solve() also has a host
//    path -- the rows in their original order on one thread, 64-bit limbs, one product per common gate (26 ns per link),
//    the values then uploaded and L, R, O gathered / checked by the same device kernel -- and B2P_SOLVE_AUTO picks by
//    a cost model of the level structure.  The device pays for wide, shallow circuits (20x a host core at 2^20 rows).
//  * Hints (gnark's hint functions: bit decompositions, inverses-or-zero, the BSB22 commitment hint, ...) are the
//    caller's: create() is told which variables a hint produces from which, places the hint at the level where its
//    inputs are known, and solve() calls the registered function there (inputs down, outputs up: a hint is a
//    synchronisation point of the device path).  Rows the caller marks unchecked (BSB22's committed rows, whose
//    qcp * pi2 term and hash the prover adds) are left out of the final gate check.  A row whose unassigned wire occurs
//    twice, or that has two unassigned wires no hint produces, is refused at create().
#pragma once
#include <algorithm>
#include <chrono>
#include <cstring>
#include <functional>
#include <vector>

#include "common.cuh"
#include "pairing_host.hpp"

namespace b2p {

constexpr uint32_t SOLVE_NONE = 0, SOLVE_O = 1, SOLVE_A = 2, SOLVE_B = 3;
constexpr uint32_t SOLVER_NARROW_MAX = 256;      // levels up to this width (one row per thread) run inside the single-block kernel
constexpr int SOLVER_NARROW_THREADS = 256;

template <class F>
struct SolverCols {
    const F *ql, *qr, *qm, *qo, *qk, *ninv_qo;   // n rows each; ninv_qo = -1/qo where the row solves its O wire
    const uint32_t *xa, *xb, *xc;
};

// one row: the unassigned wire from the assigned ones.  false: division by zero (the row cannot determine its wire)
template <class F>
__device__ __forceinline__ bool solve_row(const SolverCols<F>& c, F* values, uint32_t row, uint32_t kind) {
    const uint32_t ia = c.xa[row], ib = c.xb[row], ic = c.xc[row];
    const F ql = ld_field(c.ql + row), qr = ld_field(c.qr + row), qm = ld_field(c.qm + row), qk = ld_field(c.qk + row);
    if (kind == SOLVE_O) {
        const F a = ld_field(values + ia), b = ld_field(values + ib);
        F t = qk;
        if (!ql.is_zero()) t = t + ql * a;
        if (!qr.is_zero()) t = t + qr * b;
        if (!qm.is_zero()) t = t + qm * (a * b);
        st_field(values + ic, t * ld_field(c.ninv_qo + row));
        return true;
    }
    const F qo = ld_field(c.qo + row);
    const F other = ld_field(values + (kind == SOLVE_A ? ib : ia));
    F num = qk + (kind == SOLVE_A ? qr : ql) * other;
    if (!qo.is_zero()) num = num + qo * ld_field(values + ic);
    const F den = (kind == SOLVE_A ? ql : qr) + qm * other;
    if (den.is_zero()) return false;
    st_field(values + (kind == SOLVE_A ? ia : ib), (num * den.inverse()).neg());
    return true;
}

// ops[i] = row | kind << 30 (n <= 2^30), sorted by level
template <class F>
__global__ void __launch_bounds__(128) k_solve_level(SolverCols<F> c, F* values, const uint32_t* __restrict__ ops,
                                                     uint32_t first, uint32_t count, uint32_t* first_bad) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint32_t op = ops[first + t];
    if (!solve_row(c, values, op & 0x3FFFFFFFu, op >> 30)) atomicMin(first_bad, op & 0x3FFFFFFFu);
}
// levels [l0, l0 + nl): one block steps through them; level_off[l] = first op of level l
template <class F>
__global__ void __launch_bounds__(SOLVER_NARROW_THREADS) k_solve_narrow(SolverCols<F> c, F* values,
                                                                          const uint32_t* __restrict__ ops,
                                                                          const uint32_t* __restrict__ level_off,
                                                                          uint32_t l0, uint32_t nl, uint32_t* first_bad) {
#pragma unroll 1
    for (uint32_t l = l0; l < l0 + nl; l++) {
        const uint32_t a = level_off[l], b = level_off[l + 1];
#pragma unroll 1
        for (uint32_t i = a + threadIdx.x; i < b; i += SOLVER_NARROW_THREADS) {
            const uint32_t op = ops[i];
            if (!solve_row(c, values, op & 0x3FFFFFFFu, op >> 30)) atomicMin(first_bad, op & 0x3FFFFFFFu);
        }
        __syncthreads();
    }
}
// the caller's assignment: values[ids[i]] = in[i]
template <class F>
__global__ void k_solver_inputs(F* values, const F* __restrict__ in, const uint32_t* __restrict__ ids, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) st_field(values + ids[i], ld_field(in + i));
}
// -1 / qo for the rows that solve their O wire (once, at create)
template <class F>
__global__ void k_solver_ninv(F* ninv, const F* __restrict__ qo, const uint8_t* __restrict__ kind, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F r = F::zero();
    if (kind[i] == SOLVE_O) {
        const F q = ld_field(qo + i);
        r = (q.neg() == F::one()) ? F::one() : q.inverse().neg();
    }
    st_field(ninv + i, r);
}
// L, R, O of every row and the gate check; public rows: the public value enters through qk (gnark's completeQk)
template <class F>
__global__ void k_solver_gather(SolverCols<F> c, const F* __restrict__ values, F* L, F* R, F* O, uint64_t n,
                                uint32_t nb_public, uint32_t* first_unsat, const uint8_t* __restrict__ unchecked) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const F a = ld_field(values + c.xa[i]), b = ld_field(values + c.xb[i]), o = ld_field(values + c.xc[i]);
    st_field(L + i, a);
    st_field(R + i, b);
    st_field(O + i, o);
    if (unchecked && unchecked[i]) return;
    F t = ld_field(c.ql + i) * a + ld_field(c.qr + i) * b + ld_field(c.qm + i) * (a * b) + ld_field(c.qo + i) * o;
    t = t + (i < nb_public ? a : ld_field(c.qk + i));
    if (!t.is_zero()) atomicMin(first_unsat, (uint32_t)i);
}

struct SolverLaunch {
    uint32_t first_level, levels, first_op, ops;
    bool narrow;
    uint32_t hint_first = 0, hint_count = 0;   // hint_count > 0: not a launch but "call hints [hint_first, +count) of hint_order"
};
struct SolverHint {
    uint32_t id;
    std::vector<uint32_t> in, out;
    uint32_t level = 0;          // 1 + deepest input
    uint32_t trigger_op = 0;     // host path: runs before this entry of h_rows
};

template <class Fr>
struct Solver : SolverBase {
    using HF = hp::Fe<typename Fr::Params>;
    static_assert(sizeof(HF) == sizeof(Fr), "host and device scalars share one memory layout");

    uint64_t n = 0, nb_variables = 0;
    uint32_t nb_public = 0, nb_inputs = 0;
    cudaStream_t st = nullptr;
    // host copies (the host path and the analysis)
    std::vector<HF> h_cols[5], h_ninv;
    std::vector<uint32_t> h_x[3], h_inputs;
    std::vector<uint8_t> h_kind;
    std::vector<uint32_t> h_ops, h_level_off;
    std::vector<SolverLaunch> plan;
    uint32_t depth = 0, widest = 0;
    double est_host_us = 0, est_dev_us = 0, last_ms = 0;
    int last_where = 0;
    // device
    DevBuf<Fr> d_cols[5], d_ninv, d_values, dL, dR, dO, d_in;
    DevBuf<uint32_t> d_x[3], d_ops, d_level_off, d_flags, d_inputs;
    DevBuf<uint8_t> d_kind;
    std::vector<SolverHint> hints;
    std::vector<uint32_t> hint_order;  // hints sorted by level (device path) -- and by trigger row inside a level
    std::vector<uint32_t> hint_by_trigger;   // the same hints in the order the host path meets them
    b2p_hint_fn hint_fn = nullptr;
    void* hint_ctx = nullptr;
    DevBuf<uint8_t> d_unchecked;
    bool has_unchecked = false;
    std::vector<HF> hint_in, hint_out;
    HF* h_pinned = nullptr;            // the host path's variable vector, page-locked (nb_variables entries)
    std::vector<uint32_t> h_rows;      // the solving rows in their original order ...
    std::vector<uint8_t> h_cls;        // ... and their class: kind in the low two bits, then which coefficients matter
    static constexpr uint8_t CLS_KIND = 3, CLS_QL = 4, CLS_QR = 8, CLS_QM = 16, CLS_QM_ONE = 32, CLS_QK = 64, CLS_NINV_ONE = 128;

    ~Solver() override {
        if (graph) cudaGraphExecDestroy(graph);
        if (h_pinned) cudaFreeHost(h_pinned);
        if (st) cudaStreamDestroy(st);
    }

    void create(uint64_t n_, uint32_t nb_public_, uint64_t nb_variables_, const uint32_t* input_ids, uint32_t nb_inputs_,
                const void* const cols[5], const uint32_t* xa, const uint32_t* xb, const uint32_t* xc,
                const b2p_hint* hint_list, uint32_t n_hints, const uint8_t* unchecked) {
        n = n_; nb_public = nb_public_; nb_variables = nb_variables_; nb_inputs = nb_inputs_;
        B2P_REQUIRE(n >= 1 && n <= (1ull << 30), "solver: at most 2^30 rows");
        B2P_REQUIRE(nb_variables >= 1 && nb_variables < (1ull << 32), "solver: variable count out of range");
        B2P_REQUIRE(nb_public <= n && nb_public <= nb_inputs && nb_inputs <= nb_variables, "solver: input counts out of range");
        for (int k = 0; k < 5; k++) {
            h_cols[k].resize(n);
            memcpy(h_cols[k].data(), cols[k], n * sizeof(HF));
        }
        const uint32_t* xs[3] = {xa, xb, xc};
        for (int k = 0; k < 3; k++) {
            h_x[k].assign(xs[k], xs[k] + n);
            for (uint32_t v : h_x[k]) B2P_REQUIRE(v < nb_variables, "solver: wire refers to a variable that does not exist");
        }
        h_inputs.assign(input_ids, input_ids + nb_inputs);
        // ---- analysis: which wire a row solves, and at which level
        constexpr uint32_t UNKNOWN = 0xFFFFFFFFu;
        std::vector<uint32_t> var_level(nb_variables, UNKNOWN);
        for (uint32_t v : h_inputs) {
            B2P_REQUIRE(v < nb_variables, "solver: input id out of range");
            B2P_REQUIRE(var_level[v] == UNKNOWN, "solver: an input variable is listed twice");
            var_level[v] = 0;
        }
        for (uint32_t i = 0; i < nb_public; i++)
            B2P_REQUIRE(var_level[h_x[0][i]] == 0, "solver: a public row's L wire is not an input");
        h_kind.assign(n, SOLVE_NONE);
        std::vector<uint32_t> row_level(n, 0);
        // hints: which variable comes out of which hint; a hint is placed when a row first needs one of its outputs
        std::vector<int32_t> hint_of_var(n_hints ? nb_variables : 0, -1);
        std::vector<uint8_t> hint_placed(n_hints, 0);
        std::vector<uint64_t> hint_trigger_row(n_hints, 0);
        hints.resize(n_hints);
        for (uint32_t h = 0; h < n_hints; h++) {
            B2P_REQUIRE(hint_list[h].n_out >= 1 && (hint_list[h].in_vars || hint_list[h].n_in == 0) && hint_list[h].out_vars,
                        "solver: a hint needs outputs");
            hints[h].id = hint_list[h].id;
            hints[h].in.assign(hint_list[h].in_vars, hint_list[h].in_vars + hint_list[h].n_in);
            hints[h].out.assign(hint_list[h].out_vars, hint_list[h].out_vars + hint_list[h].n_out);
            for (uint32_t v : hints[h].in) B2P_REQUIRE(v < nb_variables, "solver: hint input out of range");
            for (uint32_t v : hints[h].out) {
                B2P_REQUIRE(v < nb_variables, "solver: hint output out of range");
                B2P_REQUIRE(var_level[v] == UNKNOWN && hint_of_var[v] < 0, "solver: a hint output is an input or another hint's output");
                hint_of_var[v] = (int32_t)h;
            }
        }
        std::function<void(uint32_t, uint64_t)> place_hint = [&](uint32_t var, uint64_t row) {
            if (hint_of_var.empty() || hint_of_var[var] < 0 || hint_placed[hint_of_var[var]]) return;
            const uint32_t h = (uint32_t)hint_of_var[var];
            hint_placed[h] = 1;                          // (also ends a cycle of hints feeding each other)
            uint32_t lvl = 0;
            for (uint32_t v : hints[h].in) {
                if (var_level[v] == UNKNOWN) place_hint(v, row);     // an input that is itself a hint's output
                if (var_level[v] == UNKNOWN)
                    throw Error(B2P_ERR_ARG, "solver: hint " + std::to_string(h) + " is needed at row " + std::to_string(row) +
                                                 " before its input variable " + std::to_string(v) + " is assigned");
                lvl = std::max(lvl, var_level[v]);
            }
            hints[h].level = lvl + 1;
            for (uint32_t v : hints[h].out) var_level[v] = lvl + 1;
            hint_placed[h] = 1;
            hint_trigger_row[h] = row;
            hint_by_trigger.push_back(h);
            depth = std::max(depth, lvl + 1);
        };
        for (uint64_t i = nb_public; i < n; i++) {
            const bool use[3] = {!h_cols[0][i].is_zero() || !h_cols[2][i].is_zero(),
                                 !h_cols[1][i].is_zero() || !h_cols[2][i].is_zero(), !h_cols[3][i].is_zero()};
            const uint32_t w[3] = {h_x[0][i], h_x[1][i], h_x[2][i]};
            for (int k = 0; k < 3; k++)
                if (use[k] && var_level[w[k]] == UNKNOWN) place_hint(w[k], i);
            uint32_t unknown_var = UNKNOWN, lvl = 0;
            int unknown_pos = -1, occurrences = 0;
            bool two = false;
            for (int k = 0; k < 3; k++) {
                if (!use[k]) continue;
                if (var_level[w[k]] == UNKNOWN) {
                    if (unknown_var != UNKNOWN && unknown_var != w[k]) two = true;
                    unknown_var = w[k];
                    unknown_pos = k;
                    occurrences++;
                } else {
                    lvl = std::max(lvl, var_level[w[k]]);
                }
            }
            if (unknown_var == UNKNOWN) continue;            // an assertion: checked with every other row at the end
            if (two)
                throw Error(B2P_ERR_ARG, "solver: row " + std::to_string(i) + " has two unassigned wires (a hint would be needed)");
            if (occurrences > 1)
                throw Error(B2P_ERR_ARG, "solver: row " + std::to_string(i) + " uses its unassigned wire twice (not linear in it)");
            h_kind[i] = unknown_pos == 2 ? SOLVE_O : unknown_pos == 0 ? SOLVE_A : SOLVE_B;
            row_level[i] = lvl + 1;
            var_level[unknown_var] = lvl + 1;
            depth = std::max(depth, lvl + 1);
        }
        for (uint64_t i = 0; i < n; i++)
            for (int k = 0; k < 3; k++) {
                if (var_level[h_x[k][i]] == UNKNOWN) place_hint(h_x[k][i], n);     // only ever read through a zero selector
                if (var_level[h_x[k][i]] == UNKNOWN)
                    throw Error(B2P_ERR_ARG, "solver: variable " + std::to_string(h_x[k][i]) + " (row " + std::to_string(i) +
                                                 ") is neither an input nor determined by a row");
            }
        // ---- rows by level (counting sort keeps the row order inside a level)
        h_level_off.assign(depth + 2, 0);
        for (uint64_t i = 0; i < n; i++)
            if (h_kind[i]) h_level_off[row_level[i] + 1]++;
        for (uint32_t l = 1; l < depth + 2; l++) h_level_off[l] += h_level_off[l - 1];
        h_ops.resize(h_level_off[depth + 1]);
        {
            std::vector<uint32_t> cur(h_level_off.begin(), h_level_off.end() - 1);
            for (uint64_t i = 0; i < n; i++)
                if (h_kind[i]) h_ops[cur[row_level[i]]++] = (uint32_t)i | ((uint32_t)h_kind[i] << 30);
        }
        // ---- launch plan + cost model (us): a wide level = one launch, a run of narrow levels = one block; the hints
        // of level l run after the rows of the levels below it and before anything that can read their outputs
        est_dev_us = 0;
        auto emit_rows = [&](uint32_t l0, uint32_t l1) {            // levels [l0, l1)
            for (uint32_t l = l0; l < l1;) {
                const uint32_t w = h_level_off[l + 1] - h_level_off[l];
                widest = std::max(widest, w);
                if (w > SOLVER_NARROW_MAX) {
                    plan.push_back({l, 1, h_level_off[l], w, false});
                    est_dev_us += 4.0 + w * 2.5e-4;
                    l++;
                    continue;
                }
                uint32_t e = l;
                while (e < l1 && h_level_off[e + 1] - h_level_off[e] <= SOLVER_NARROW_MAX) {
                    widest = std::max(widest, h_level_off[e + 1] - h_level_off[e]);
                    e++;
                }
                plan.push_back({l, e - l, h_level_off[l], h_level_off[e] - h_level_off[l], true});
                est_dev_us += 4.0 + (e - l) * 2.2;
                l = e;
            }
        };
        for (uint32_t h = 0; h < n_hints; h++)
            if (hint_placed[h]) hint_order.push_back(h);
        std::stable_sort(hint_order.begin(), hint_order.end(),
                         [&](uint32_t x, uint32_t y) { return hints[x].level < hints[y].level; });
        {
            uint32_t start = 1;
            for (size_t k = 0; k < hint_order.size();) {
                const uint32_t l = hints[hint_order[k]].level;
                size_t e = k;
                while (e < hint_order.size() && hints[hint_order[e]].level == l) e++;
                emit_rows(start, l);
                SolverLaunch item{l, 0, 0, 0, false};
                item.hint_first = (uint32_t)k;
                item.hint_count = (uint32_t)(e - k);
                plan.push_back(item);
                est_dev_us += 30.0 * (e - k);
                start = l;
                k = e;
            }
            emit_rows(start, depth + 1);
        }
        // host path: a hint runs before the first solving row at or after the trace row that first needed it
        {
            size_t k = 0;
            std::vector<uint32_t> solving;
            for (uint64_t i = nb_public; i < n; i++)
                if (h_kind[i]) solving.push_back((uint32_t)i);
            for (uint32_t h : hint_by_trigger) {
                while (k < solving.size() && solving[k] < hint_trigger_row[h]) k++;
                hints[h].trigger_op = (uint32_t)k;
            }
        }
        if (unchecked) {
            has_unchecked = true;
            d_unchecked.alloc(n);
        }
        est_dev_us += 10 + n * 1e-4;
        est_host_us = h_ops.size() * 0.06 + nb_variables * 0.004 + 300;   // one product per common row + upload of the values
        // ---- device copies
        B2P_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        for (int k = 0; k < 5; k++) {
            d_cols[k].alloc(n);
            B2P_CUDA(cudaMemcpyAsync(d_cols[k].p, h_cols[k].data(), n * sizeof(Fr), cudaMemcpyHostToDevice, st));
        }
        for (int k = 0; k < 3; k++) {
            d_x[k].alloc(n);
            B2P_CUDA(cudaMemcpyAsync(d_x[k].p, h_x[k].data(), n * 4, cudaMemcpyHostToDevice, st));
        }
        d_kind.alloc(n);
        B2P_CUDA(cudaMemcpyAsync(d_kind.p, h_kind.data(), n, cudaMemcpyHostToDevice, st));
        if (has_unchecked) B2P_CUDA(cudaMemcpyAsync(d_unchecked.p, unchecked, n, cudaMemcpyHostToDevice, st));
        d_ops.alloc(std::max<size_t>(h_ops.size(), 1));
        if (!h_ops.empty()) B2P_CUDA(cudaMemcpyAsync(d_ops.p, h_ops.data(), h_ops.size() * 4, cudaMemcpyHostToDevice, st));
        d_level_off.alloc(h_level_off.size());
        B2P_CUDA(cudaMemcpyAsync(d_level_off.p, h_level_off.data(), h_level_off.size() * 4, cudaMemcpyHostToDevice, st));
        d_inputs.alloc(std::max<uint32_t>(nb_inputs, 1));
        if (nb_inputs) B2P_CUDA(cudaMemcpyAsync(d_inputs.p, h_inputs.data(), nb_inputs * 4, cudaMemcpyHostToDevice, st));
        d_ninv.alloc(n);
        B2P_LAUNCH((k_solver_ninv<Fr>), div_up(n, 128), 128, 0, st, d_ninv.p, d_cols[3].p, d_kind.p, n);
        h_ninv.resize(n);
        B2P_CUDA(cudaMemcpyAsync(h_ninv.data(), d_ninv.p, n * sizeof(Fr), cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));                 // h_ninv is complete
        for (uint64_t i = nb_public; i < n; i++) {
            if (!h_kind[i]) continue;
            uint8_t cls = h_kind[i];
            if (!h_cols[0][i].is_zero()) cls |= CLS_QL;
            if (!h_cols[1][i].is_zero()) cls |= CLS_QR;
            if (!h_cols[2][i].is_zero()) cls |= CLS_QM;
            if (h_cols[2][i] == HF::one()) cls |= CLS_QM_ONE;
            if (!h_cols[4][i].is_zero()) cls |= CLS_QK;
            if (h_ninv[i] == HF::one()) cls |= CLS_NINV_ONE;
            h_rows.push_back((uint32_t)i);
            h_cls.push_back(cls);
        }
        B2P_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_pinned), nb_variables * sizeof(HF)));
        d_values.alloc(nb_variables);
        d_in.alloc(std::max<uint32_t>(nb_inputs, 1));
        dL.alloc(n); dR.alloc(n); dO.alloc(n);
        d_flags.alloc(2);
        B2P_CUDA(cudaStreamSynchronize(st));
    }

    SolverCols<Fr> dcols() const {
        return {d_cols[0].p, d_cols[1].p, d_cols[2].p, d_cols[3].p, d_cols[4].p, d_ninv.p, d_x[0].p, d_x[1].p, d_x[2].p};
    }

    void set_hint_fn(b2p_hint_fn fn, void* ctx) override { hint_fn = fn; hint_ctx = ctx; }
    void info(uint64_t* out) const override {
        out[0] = depth; out[1] = widest; out[2] = h_ops.size(); out[3] = plan.size();
        out[4] = (uint64_t)est_host_us; out[5] = (uint64_t)est_dev_us; out[6] = (uint64_t)(last_ms * 1000.0); out[7] = last_where;
    }

    int choose(int where) const {
        if (where == B2P_SOLVE_HOST || where == B2P_SOLVE_DEVICE) return where;
        B2P_REQUIRE(where == B2P_SOLVE_AUTO, "solver: unknown placement");
        return est_dev_us < est_host_us ? B2P_SOLVE_DEVICE : B2P_SOLVE_HOST;
    }

    // inputs: nb_inputs Fr (Montgomery) in the order of the input ids given at create
    void solve(const void* inputs, int where, void* L, void* R, void* O, bool device_out, void** dptrs) override {
        B2P_REQUIRE(inputs || nb_inputs == 0, "null argument");
        where = choose(where);
        last_where = where;
        const auto t0 = std::chrono::steady_clock::now();
        // either way L, R, O are gathered -- and every row checked -- by k_solver_gather on the device
        if (where == B2P_SOLVE_DEVICE) solve_device(inputs);
        else solve_host(inputs);
        if (device_out) {
            dptrs[0] = dL.p; dptrs[1] = dR.p; dptrs[2] = dO.p;
        } else {
            B2P_CUDA(cudaMemcpyAsync(L, dL.p, n * sizeof(Fr), cudaMemcpyDeviceToHost, st));
            B2P_CUDA(cudaMemcpyAsync(R, dR.p, n * sizeof(Fr), cudaMemcpyDeviceToHost, st));
            B2P_CUDA(cudaMemcpyAsync(O, dO.p, n * sizeof(Fr), cudaMemcpyDeviceToHost, st));
            B2P_CUDA(cudaStreamSynchronize(st));
        }
        last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }

    // the fixed part of a device solve: flags, input scatter, one launch per plan entry (hints: a round trip through
    // the caller's function), gather + check
    void queue_levels() {
        const SolverCols<Fr> c = dcols();
        B2P_CUDA(cudaMemsetAsync(d_flags.p, 0xFF, 2 * sizeof(uint32_t), st));
        if (nb_inputs)
            B2P_LAUNCH((k_solver_inputs<Fr>), div_up(nb_inputs, 128), 128, 0, st, d_values.p, d_in.p, d_inputs.p, nb_inputs);
        for (const SolverLaunch& s : plan) {
            if (s.hint_count) {
                for (uint32_t k = 0; k < s.hint_count; k++) run_hint_device(hints[hint_order[s.hint_first + k]]);
            } else if (s.narrow) {
                B2P_LAUNCH((k_solve_narrow<Fr>), 1, SOLVER_NARROW_THREADS, 0, st, c, d_values.p, d_ops.p, d_level_off.p,
                           s.first_level, s.levels, d_flags.p);
            } else {
                B2P_LAUNCH((k_solve_level<Fr>), div_up(s.ops, 128), 128, 0, st, c, d_values.p, d_ops.p, s.first_op, s.ops,
                           d_flags.p);
            }
        }
        B2P_LAUNCH((k_solver_gather<Fr>), div_up(n, 128), 128, 0, st, c, d_values.p, dL.p, dR.p, dO.p, n, nb_public,
                   d_flags.p + 1, has_unchecked ? d_unchecked.p : nullptr);
    }
    void call_hint(const SolverHint& h) {
        if (!hint_fn) throw Error(B2P_ERR_ARG, "solver: the circuit has hints but no hint function is set (b2p_solver_set_hint_fn)");
        if (hint_fn(hint_ctx, h.id, hint_in.data(), (uint32_t)h.in.size(), hint_out.data(), (uint32_t)h.out.size()) != 0)
            throw Error(B2P_ERR_INTERNAL, "solver: the hint function reported a failure (hint id " + std::to_string(h.id) + ")");
    }
    // inputs down, the caller's function, outputs up: a synchronisation point of the device path
    void run_hint_device(const SolverHint& h) {
        hint_in.resize(std::max<size_t>(h.in.size(), 1));
        hint_out.resize(h.out.size());
        for (size_t j = 0; j < h.in.size(); j++)
            B2P_CUDA(cudaMemcpyAsync(&hint_in[j], d_values.p + h.in[j], sizeof(Fr), cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));
        call_hint(h);
        for (size_t j = 0; j < h.out.size(); j++)
            B2P_CUDA(cudaMemcpyAsync(d_values.p + h.out[j], &hint_out[j], sizeof(Fr), cudaMemcpyHostToDevice, st));
        B2P_CUDA(cudaStreamSynchronize(st));          // hint_out is reused by the next hint
    }
    void run_hint_host(const SolverHint& h, HF* v) {
        hint_in.resize(std::max<size_t>(h.in.size(), 1));
        hint_out.resize(h.out.size());
        for (size_t j = 0; j < h.in.size(); j++) hint_in[j] = v[h.in[j]];
        call_hint(h);
        for (size_t j = 0; j < h.out.size(); j++) v[h.out[j]] = hint_out[j];
    }
    // A shallow circuit is hundreds of short launches with fixed arguments: captured once into a CUDA graph, a solve is
    // one graph launch (B2P_SOLVER_GRAPH=0 keeps the plain launches; so does any failure to capture).
    cudaGraphExec_t graph = nullptr;
    bool graph_tried = false;
    void build_graph() {
        graph_tried = true;
        const char* e = getenv("B2P_SOLVER_GRAPH");
        if ((e && atoi(e) == 0) || plan.size() < 8 || !hint_order.empty()) return;    // a hint is a host call: no capture
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return; }
        bool ok = true;
        try { queue_levels(); } catch (...) { ok = false; }
        if (cudaStreamEndCapture(st, &g) != cudaSuccess || !ok || !g) { cudaGetLastError(); if (g) cudaGraphDestroy(g); return; }
        if (cudaGraphInstantiate(&graph, g, 0) != cudaSuccess) { cudaGetLastError(); graph = nullptr; }
        cudaGraphDestroy(g);
    }
    void solve_device(const void* inputs) {
        // only the inputs travel; every other variable is written by the row that determines it (create() checked that)
        if (nb_inputs) B2P_CUDA(cudaMemcpyAsync(d_in.p, inputs, nb_inputs * sizeof(Fr), cudaMemcpyHostToDevice, st));
        if (!graph_tried) build_graph();
        if (graph) {
            B2P_CUDA(cudaGraphLaunch(graph, st));
            g_launch_count += plan.size() + 2;
        } else {
            queue_levels();
        }
        uint32_t flags[2];
        B2P_CUDA(cudaMemcpyAsync(flags, d_flags.p, sizeof flags, cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));
        report(flags[0], flags[1]);
    }

    static void report(uint32_t div0_row, uint32_t unsat_row) {
        if (div0_row != 0xFFFFFFFFu)
            throw Error(B2P_ERR_VERIFY, "solver: row " + std::to_string(div0_row) + " cannot determine its wire (division by zero)");
        if (unsat_row != 0xFFFFFFFFu)
            throw Error(B2P_ERR_VERIFY, "constraint #" + std::to_string(unsat_row) + " is not satisfied");
    }

    // The rows in their original order on one host thread (a dependency chain is the whole job: 0.05 us per link here
    // against 2 us on an SM).  Row classes computed at create() keep the common gates at one product per row: a
    // coefficient that is 0 or 1 is neither loaded nor multiplied.  The values then go up (page-locked, 32 B per
    // variable) and the device gathers L, R, O and checks every row, exactly as after a device solve.
    void solve_host(const void* inputs) {
        HF* v = h_pinned;
        const HF* in = static_cast<const HF*>(inputs);
        for (uint32_t i = 0; i < nb_inputs; i++) v[h_inputs[i]] = in[i];
        uint32_t div0 = 0xFFFFFFFFu;
        const uint32_t nops = (uint32_t)h_rows.size();
        size_t next_hint = 0;
        for (uint32_t k = 0; k < nops; k++) {
            while (next_hint < hint_by_trigger.size() && hints[hint_by_trigger[next_hint]].trigger_op <= k)
                run_hint_host(hints[hint_by_trigger[next_hint++]], v);
            const uint64_t i = h_rows[k];
            const uint8_t cls = h_cls[k];
            const uint32_t ia = h_x[0][i], ib = h_x[1][i], ic = h_x[2][i];
            if ((cls & CLS_KIND) == SOLVE_O) {
                HF t;
                if (cls & CLS_QM) {
                    t = v[ia] * v[ib];
                    if (!(cls & CLS_QM_ONE)) t = t * h_cols[2][i];
                    if (cls & CLS_QK) t = t + h_cols[4][i];
                } else {
                    t = (cls & CLS_QK) ? h_cols[4][i] : HF::zero();
                }
                if (cls & CLS_QL) t = t + h_cols[0][i] * v[ia];
                if (cls & CLS_QR) t = t + h_cols[1][i] * v[ib];
                v[ic] = (cls & CLS_NINV_ONE) ? t : t * h_ninv[i];
                continue;
            }
            const bool solve_a = (cls & CLS_KIND) == SOLVE_A;
            const HF &ql = h_cols[0][i], &qr = h_cols[1][i], &qm = h_cols[2][i], &qo = h_cols[3][i], &qk = h_cols[4][i];
            const HF other = v[solve_a ? ib : ia];
            HF num = qk + (solve_a ? qr : ql) * other;
            if (!qo.is_zero()) num = num + qo * v[ic];
            const HF den = (solve_a ? ql : qr) + qm * other;
            if (den.is_zero()) { div0 = std::min<uint32_t>(div0, (uint32_t)i); continue; }
            v[solve_a ? ia : ib] = (num * den.inverse()).neg();
        }
        while (next_hint < hint_by_trigger.size()) run_hint_host(hints[hint_by_trigger[next_hint++]], v);
        if (div0 != 0xFFFFFFFFu) report(div0, 0xFFFFFFFFu);
        B2P_CUDA(cudaMemcpyAsync(d_values.p, v, nb_variables * sizeof(Fr), cudaMemcpyHostToDevice, st));
        B2P_CUDA(cudaMemsetAsync(d_flags.p, 0xFF, 2 * sizeof(uint32_t), st));
        B2P_LAUNCH((k_solver_gather<Fr>), div_up(n, 128), 128, 0, st, dcols(), d_values.p, dL.p, dR.p, dO.p, n, nb_public,
                   d_flags.p + 1, has_unchecked ? d_unchecked.p : nullptr);
        uint32_t flags[2];
        B2P_CUDA(cudaMemcpyAsync(flags, d_flags.p, sizeof flags, cudaMemcpyDeviceToHost, st));
        B2P_CUDA(cudaStreamSynchronize(st));
        report(flags[0], flags[1]);
    }
};

}  // namespace b2p
