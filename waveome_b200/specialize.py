"""Run-time specialisation of the element-wise passes (Gram builder, gradient reduction) for ONE kernel program.

The interpreter kernels of ``csrc/wv_elem.cuh`` spend ~80 % of their instructions on leaf dispatch, categorical selects
and masking (ncu, profiles/r01i_ncu_set_full_summary.txt).  Here the flat program (sum of products of leaves — the trees
of waveome/regularization.py:14-189 ``full_kernel_build`` and of the search expansions) becomes straight-line CUDA text:

* leaf types, covariate columns, slots and the trainable / frozen split are constants of the text;
* covariates are staged per tile already scaled per leaf (``x * sqrt(log2 e / 2) / lengthscale`` for squared
  exponentials), categorical columns as int32 codes (``tf.round`` == ``rint``, waveome/kernels.py:113-114), so a
  categorical leaf is an integer compare and a predicated add;
* the squared-exponential leaves of a product share ONE ``2^u``; in the Gram pass ``u`` also carries log2 of the product
  of the component's variances;
* every gradient sum of a component is ``sum_e G_e q_e`` with ``G = W o mask o prod(unit-variance values)`` formed once
  per component (q = 1 for all variances of the component, (s d)^2 for a squared-exponential lengthscale, ...); the
  scalar factors (other variances, 2 ln2 / lengthscale, ...) are applied after the reduction.

``generate(program)`` returns the CUDA source (prelude = csrc/wv_common.cuh + wv_kernels.cuh + wv_spec.cuh, then the two
kernels), or ``None`` when the program uses a leaf the generator does not cover (polynomial, empty) — the interpreter
then stays in charge.  The engine compiles the text with NVRTC for sm_100a (``Batch.specialize``).
"""
from __future__ import annotations

import hashlib
import os
import re
from dataclasses import dataclass
from typing import Dict, List, Optional

from .program import Program

SE, M12, M32, M52, PERIODIC, LINEAR, CONST, CAT, POLY, EMPTY = range(10)
_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
_PRELUDE_FILES = ("wv_common.cuh", "wv_kernels.cuh", "wv_spec.cuh")
_prelude_cache: Optional[str] = None

#: resident CTAs per SM the register allocation of the generated kernels aims for (measured on config 3, profiles/r02*)
GRAM_MIN_CTAS = int(os.environ.get("WV_SPEC_GRAM_MINB", "4"))
#: per-component run-time mask checks (feature-importance batches switch components off).  Measured on config 3: WITHOUT
#: them the kernels are slower (gram 3.28 vs 2.76 ms, grad 3.60 vs 3.37 ms per 2000-model evaluation): the uniform
#: branches keep the compiler from interleaving components, which costs registers (gram spills at 64).  Kept on; the
#: switch exists for that experiment only (a text generated without them ignores the masks).
USE_CMASK = os.environ.get("WV_SPEC_CMASK", "1") == "1"
GRAD_MIN_CTAS = int(os.environ.get("WV_SPEC_GRAD_MINB", "2"))     # 3 CTAs (80 registers) spill the gradient sums: 4.6 ms vs 3.4 ms

SE_SCALE = "0.84932180028801907"        # sqrt(log2(e) / 2): exp(-r2 / 2) = 2^(-(s (x - x'))^2), s = SE_SCALE / lengthscale
TWO_LN2 = "1.3862943611198906"          # r2 = 2 ln2 (s d)^2


def prelude() -> str:
    """The engine's own device headers, concatenated for NVRTC (local includes and include guards stripped)."""
    global _prelude_cache
    if _prelude_cache is None:
        parts = []
        for name in _PRELUDE_FILES:
            with open(os.path.join(_CSRC, name)) as fh:
                text = fh.read()
            text = re.sub(r'^\s*#\s*include\s+"[^"]+"\s*$', "", text, flags=re.M)
            text = re.sub(r"^\s*#\s*pragma\s+once\s*$", "", text, flags=re.M)
            parts.append(f"// ---- {name}\n{text}")
        _prelude_cache = "\n".join(parts)
    return _prelude_cache


@dataclass
class SpecSource:
    source: str
    gram_name: str
    grad_name: str
    gram_smem: int
    grad_smem: int
    n_sums: int
    key: str            # hash of the text: the engine's compile cache key


class _Comp:
    def __init__(self):
        self.cats: List[int] = []          # categorical array ids
        self.var_slots: List[int] = []     # variance slot of every leaf of the product
        self.se: List[tuple] = []          # (scaled array id, ls slot)
        self.other: List[dict] = []        # matern / periodic / linear leaves
        self.sums: Dict[str, int] = {}     # name -> index into the reduced sums


def _analyse(p: Program):
    arrays: Dict[tuple, int] = {}     # ("se" | "inv" | "raw", dim, slot) -> array id
    cats: Dict[int, int] = {}         # dim -> categorical array id

    def arr(kind, dim, slot=-1):
        key = (kind, int(dim), int(slot))
        return arrays.setdefault(key, len(arrays))

    comps: List[_Comp] = []
    for c in range(p.n_comp):
        cp = _Comp()
        for l in range(int(p.comp_start[c]), int(p.comp_start[c + 1])):
            t, dim = int(p.leaf_type[l]), int(p.leaf_dim[l])
            sv, sl, sa = int(p.leaf_s_var[l]), int(p.leaf_s_ls[l]), int(p.leaf_s_aux[l])
            if t in (POLY, EMPTY) or sv < 0:
                return None
            cp.var_slots.append(sv)
            if t == CAT:
                cp.cats.append(cats.setdefault(dim, len(cats)))
            elif t == CONST:
                pass
            elif t == SE:
                cp.se.append((arr("se", dim, sl), sl))
            elif t in (M12, M32, M52):
                cp.other.append(dict(type=t, arr=arr("inv", dim, sl), ls=sl, aux=-1))
            elif t == PERIODIC:
                cp.other.append(dict(type=t, arr=arr("raw", dim), ls=sl, aux=sa))
            elif t == LINEAR:
                cp.other.append(dict(type=t, arr=arr("raw", dim), ls=-1, aux=-1))
            else:
                return None
        if int(p.comp_start[c + 1]) == int(p.comp_start[c]):
            return None
        comps.append(cp)
    return arrays, cats, comps


def _prod(terms: List[str]) -> str:
    return " * ".join(terms) if terms else "1.0"


#: Gram pass: add a masked-out value as ``+ 0.0`` through a select instead of branching around the addition (the sums
#: start at +0.0 and every value is positive, so the bits are the same).  WV_SPEC_MASK_SELECT=0 restores the branch.
MASK_SELECT = os.environ.get("WV_SPEC_MASK_SELECT", "1") != "0"


def _mask_expr(cp: _Comp, b: str) -> str:
    return " && ".join(f"kr{k} == kc{k}[{b}]" for k in cp.cats)


def _emit_loads(cp: _Comp, ind: str) -> List[str]:
    """covariates of one micro-tile ROW (row r of the tile, columns c_off .. c_off + 3) for the leaves of a component;
    the categorical codes are loaded once per row for all components (see ``cat_loads``)"""
    out = []
    ids = [a for a, _ in cp.se] + [o["arr"] for o in cp.other]
    for a in dict.fromkeys(ids):
        out.append(f"{ind}const double xi{a} = sw.a[{a}][r]; double xj{a}[4]; wvs_ld4(&sw.a[{a}][c_off], xj{a});")
    return out


def _warp_skip_open(cp: _Comp, ind: str) -> List[str]:
    """whole warps skip the transcendental factors of a product whose categorical mask is zero for the warp's row group"""
    any_ = " || ".join(f"({_mask_expr(cp, str(b))})" for b in range(4))
    return [f"{ind}if (__any_sync(0xffffffffu, {any_})) {{"]


def generate(p: Program) -> Optional[SpecSource]:
    an = _analyse(p)
    if an is None:
        return None
    arrays, cats, comps = an
    ns, nc = int(p.n_slots), len(comps)
    trainable = [int(x) >= 0 for x in p.slot_xindex]
    na, ncat = max(1, len(arrays)), max(1, len(cats))

    # ---- constants staged once per CTA: scale of every staged array, per component log2(prod var) or prod var
    kc_expr: List[str] = []
    arr_scale: Dict[int, int] = {}
    for (kind, dim, slot), a in arrays.items():
        if kind == "se":
            arr_scale[a] = len(kc_expr); kc_expr.append(f"{SE_SCALE} / sm.theta[{slot}]")
        elif kind == "inv":
            arr_scale[a] = len(kc_expr); kc_expr.append(f"1.0 / sm.theta[{slot}]")
    comp_kc: List[int] = []
    for cp in comps:
        var_all = _prod([f"sm.theta[{s}]" for s in cp.var_slots])
        comp_kc.append(len(kc_expr))
        kc_expr.append(f"log2({var_all})" if cp.se else var_all)
    nkc = max(1, len(kc_expr))

    # ---- gradient sums: per component S (all its variances) + one per lengthscale-like parameter
    nsum = 0
    for cp in comps:
        cp.sums = {}
        needs_S = any(trainable[s] for s in cp.var_slots)
        for i, (_, sl) in enumerate(cp.se):
            if trainable[sl]:
                cp.sums[f"se{i}"] = -1
        for i, o in enumerate(cp.other):
            if o["ls"] >= 0 and trainable[o["ls"]]:
                cp.sums[f"ls{i}"] = -1
            if o["aux"] >= 0 and trainable[o["aux"]]:
                cp.sums[f"aux{i}"] = -1
        if needs_S or cp.sums:
            if needs_S:
                cp.sums = {"S": -1, **cp.sums}
            for k in cp.sums:
                cp.sums[k] = nsum
                nsum += 1
    trw_sum = nsum
    nsum += 1

    body: List[str] = []
    w = body.append
    # the text depends on the structure only (leaves, slots, trainable / frozen split) -- not on priors, transforms or
    # frozen values, which the kernels read from the device program: programs that differ in those share one compilation
    tag = hashlib.sha1(repr((p.comp_start.tolist(), p.leaf_type.tolist(), p.leaf_dim.tolist(), p.leaf_s_var.tolist(),
                             p.leaf_s_ls.tolist(), p.leaf_s_aux.tolist(), trainable, int(p.noise_slot), int(p.mean_slot),
                             ns)).encode()).hexdigest()[:12]
    gram_name, grad_name = f"wvs_gram_{tag}", f"wvs_grad_{tag}"

    red_doubles = nsum * 33
    stage_bytes = 8 * 48 * na + 4 * 48 * ncat                 # one warp's staged covariates
    warp_union = max(stage_bytes, 8 * red_doubles)
    warp_union += (-warp_union) % 16

    def smem_struct(name, with_red):
        """shared memory of one CTA; returns its size in bytes"""
        w(f"struct {name}Warp {{")
        if with_red:
            w("  union {")
            w(f"    struct {{ double a[{na}][WVS_STAGE]; int c[{ncat}][WVS_STAGE]; }};")
            w(f"    double red[{warp_union // 8}];      // the lanes' sums, after the rows are done with the staged columns")
            w("  };")
            w("  double al[WVS_STAGE];")
            per_warp = warp_union + 8 * 48
        else:
            w(f"  double a[{na}][WVS_STAGE];")
            w(f"  int c[{ncat}][WVS_STAGE];")
            per_warp = stage_bytes + (-stage_bytes) % 16
            if per_warp != stage_bytes:
                w(f"  int pad_[{(per_warp - stage_bytes) // 4}];")
        w("};")
        w(f"struct {name} {{")
        w(f"  double theta[{ns + (ns & 1)}];")
        w(f"  double kc[{nkc + (nkc & 1)}];")
        w("  double tab[WV_EXP2_BIG_TAB];")
        w(f"  {name}Warp w[WVS_THREADS / 32];")
        w("  int next_unit, pad_unit;         // work counter: (tile, region) units are pulled by whichever warp is free")
        size = 8 * (ns + (ns & 1)) + 8 * (nkc + (nkc & 1)) + 8 * 2048 + 8 * per_warp + 8
        if with_red:
            w(f"  double wsum[WVS_TPC_MAX][8][{nsum}];    // per tile and region, combined in fixed order")
            w(f"  double sums[WVS_TPC_MAX][{nsum}];")
            size += 8 * 8 * 8 * nsum + 8 * 8 * nsum
        w("};")
        w(f"static_assert(sizeof({name}) == {size}, \"shared-memory layout\");")
        return size

    def stage_model():
        w("  const int b = active[blockIdx.y];")
        w(f"  wvs_stage_theta(bd, b, xall, sm.theta, {ns});")
        w("  wvs_load_tab(sm.tab, gtab);")
        if USE_CMASK:
            w("  const unsigned cmask = bd.comp_mask[b];")
        w("  __syncthreads();")
        for i, e in enumerate(kc_expr):
            w(f"  if (threadIdx.x == {i % 256}) sm.kc[{i}] = {e};")
        w("  if (threadIdx.x == 255) sm.next_unit = 0;")
        w("  __syncthreads();")
        w("  const int n = bd.n, ld = bd.npad;")
        w("  const int t0 = blockIdx.x * tpc, t1 = min(ntiles, t0 + tpc);")
        w("  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;")
        w("  // Work units = (tile, 16 x 32 region) pairs, pulled from a shared counter: warps whose regions are masked out")
        w("  // (categorical x numeric products) or lie above the diagonal finish early and take the next unit instead of")
        w("  // idling until the slowest warp of the CTA is done (13 % of the stall samples with a fixed warp -> region map).")
        w("  // Results are indexed by unit, not by warp: sums stay bit-reproducible.")
        w("  const int n_units = (t1 - t0) * 8;")
        w("  const int r_loc = (lane >> 3) * 4, c_off = 16 + (lane & 7) * 4;  // the lane's rows / columns in the staged arrays")
        w("  auto& sw = sm.w[warp];")

    def stage_warp(with_alpha):
        w("    __syncwarp();       // the previous tile's rows are done with the staged columns")
        w("    for (int i = lane; i < WVS_STAGE; i += 32) {")
        w("      const size_t g = i < 16 ? (size_t)ti * 64 + wr + i : (size_t)tj * 64 + wc + (i - 16);")
        dims = sorted({d for (_, d, _) in arrays} | set(cats))
        for d in dims:
            w(f"      const double x{d} = bd.Xt[(size_t){d} * ld + g];")
        for (kind, dim, slot), a in arrays.items():
            sc = f" * sm.kc[{arr_scale[a]}]" if a in arr_scale else ""
            w(f"      sw.a[{a}][i] = x{dim}{sc};")
        for dim, k in cats.items():
            w(f"      sw.c[{k}][i] = (int)rint(x{dim});")
        if with_alpha:
            w("      sw.al[i] = al[g];")
        w("    }")
        w("    __syncwarp();")

    def value_terms(cp: _Comp, a: str, b: str, grads: bool) -> List[str]:
        """statements that multiply the unit-variance values of the non-SE leaves into `v` (and set their q's)"""
        out = []
        for i, o in enumerate(cp.other):
            x = o["arr"]
            if o["type"] == LINEAR:
                out.append(f"v *= xi{x} * xj{x}[{b}];")
            elif o["type"] in (M12, M32, M52):
                if grads and f"ls{i}" in cp.sums:
                    out.append(f"double E{i}, q{i}; wvs_matern<{o['type']}>(xi{x}, xj{x}[{b}], E{i}, q{i}); v *= E{i};")
                else:
                    out.append(f"v *= wvs_matern_value<{o['type']}>(xi{x}, xj{x}[{b}]);")
            elif o["type"] == PERIODIC:
                if grads and (f"ls{i}" in cp.sums or f"aux{i}" in cp.sums):
                    out.append(f"double E{i}, q{i}, qa{i}; wvs_periodic(xi{x}, xj{x}[{b}], sm.theta[{o['ls']}], "
                               f"sm.theta[{o['aux']}], E{i}, q{i}, qa{i}); v *= E{i};")
                else:
                    out.append(f"v *= wvs_periodic_value(xi{x}, xj{x}[{b}], sm.theta[{o['ls']}], sm.theta[{o['aux']}]);")
        return out

    def cat_loads(ind):
        for k in range(len(cats)):
            w(f"{ind}const int kr{k} = sw.c[{k}][r]; int kc{k}[4]; wvs_ld4i(&sw.c[{k}][c_off], kc{k});")

    # The micro-tile is walked ROW BY ROW in a rolled loop: the text of one row (4 elements per component) is a quarter
    # of the fully unrolled micro-tile -- the unrolled kernels stalled on instruction fetch (22 % of the warp stall
    # samples, profiles/r02b) -- and the live state is one row, so four CTAs fit on an SM instead of two.
    # =============================================================================== gram
    w("// ---------------- generated: Gram builder")
    gram_smem = smem_struct("WvsGramSmem", False)
    w(f"extern \"C\" __global__ void __launch_bounds__(WVS_THREADS, {GRAM_MIN_CTAS}) {gram_name}(WvBatchDev bd, const int* __restrict__ active,")
    w("    const double* __restrict__ xall, const double* __restrict__ gtab, int ntiles, int tpc) {")
    w("  WvsGramSmem& sm = *reinterpret_cast<WvsGramSmem*>(wvs_smem_raw);")
    stage_model()
    w("  double* Ab = bd.A + (size_t)b * ld * ld;")
    w("  const double* yb = bd.Y + (size_t)b * ld;")
    w("  const double* lam = bd.site_lam ? bd.site_lam + (size_t)b * ld : nullptr;")
    w("  const double* eta = bd.site_eta ? bd.site_eta + (size_t)b * ld : nullptr;")
    w("  for (;;) {")
    w("    int unit = 0;")
    w("    if (lane == 0) unit = atomicAdd(&sm.next_unit, 1);")
    w("    unit = __shfl_sync(0xffffffffu, unit, 0);")
    w("    if (unit >= n_units) break;")
    w("    const int t = t0 + (unit >> 3), reg = unit & 7;")
    w("    const int wr = (reg >> 1) * 16, wc = (reg & 1) * 32;          // the region inside tile t")
    w("    int ti, tj;")
    w("    wv_tile_from_linear(t, ti, tj);")
    w("    if (ti == tj && wc > wr + 15) continue;      // the strict upper part of a diagonal tile is never read")
    stage_warp(False)
    w("    const double s2 = sm.theta[%d];" % int(p.noise_slot))
    w("    const double cmean = %s;" % (f"sm.theta[{int(p.mean_slot)}]" if int(p.mean_slot) >= 0 else "0.0"))
    w("#pragma unroll 1")
    w("    for (int a = 0; a < 4; ++a) {")
    w("      const int r = r_loc + a;")
    w("      double acc[4] = {0.0, 0.0, 0.0, 0.0};")
    cat_loads("      ")
    for c, cp in enumerate(comps):
        expensive = bool(cp.se or cp.other)
        w(f"      if (cmask & {1 << c}u) {{      // component {c}" if USE_CMASK else f"      {{      // component {c}")
        skip = bool(cp.cats) and expensive
        ind = "        "
        if skip:
            for s_ in _warp_skip_open(cp, ind):
                w(s_)
            ind += "  "
        for s_ in _emit_loads(cp, ind):
            w(s_)
        w(f"{ind}const double k_ = sm.kc[{comp_kc[c]}];")
        w("#pragma unroll")
        w(f"{ind}for (int j = 0; j < 4; ++j) {{")
        if cp.se:
            w(f"{ind}  double u = k_;")
            for (x, _) in cp.se:
                w(f"{ind}  {{ const double d = xi{x} - xj{x}[j]; u = fma(-d, d, u); }}")
            w(f"{ind}  double v = wv_exp2_big_lo(u, sm.tab);")
        else:
            w(f"{ind}  double v = k_;")
        for s_ in value_terms(cp, "a", "j", False):
            w(f"{ind}  " + s_)
        if cp.cats:
            w(f"{ind}  acc[j] += ({_mask_expr(cp, 'j')}) ? v : 0.0;" if MASK_SELECT else f"{ind}  if ({_mask_expr(cp, 'j')}) acc[j] += v;")
        else:
            w(f"{ind}  acc[j] += v;")
        w(f"{ind}}}")
        if skip:
            w("        }")
        w("      }")
    w("""      const int gi = ti * 64 + wr + r, gj0 = tj * 64 + wc + (c_off - 16);
      double out[4];
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int gj = gj0 + bb;
        double v;
        if (gi < n && gj < n) v = acc[bb] + (gi == gj ? (lam ? bd.jitter + 1.0 / lam[gi] : s2) : 0.0);
        else if (gi == n && gj < n) v = (lam ? eta[gj] / lam[gj] : yb[gj]) - cmean;     // RHS row d^T
        else v = (gi == gj) ? 1.0 : 0.0;                     // identity padding (incl. A[n][n] = 1)
        out[bb] = v;
      }
      double2* dst = reinterpret_cast<double2*>(Ab + (size_t)gi * ld + gj0);
      dst[0] = make_double2(out[0], out[1]);
      dst[1] = make_double2(out[2], out[3]);
    }
  }
}
""")

    # =============================================================================== grad
    w("// ---------------- generated: gradient reduction")
    grad_smem = smem_struct("WvsGradSmem", True)
    w(f"extern \"C\" __global__ void __launch_bounds__(WVS_THREADS, {GRAD_MIN_CTAS}) {grad_name}(WvBatchDev bd, const int* __restrict__ active,")
    w("    const double* __restrict__ xall, const double* __restrict__ gtab, int ntiles, int tpc) {")
    w("  WvsGradSmem& sm = *reinterpret_cast<WvsGradSmem*>(wvs_smem_raw);")
    stage_model()
    w("  const double* Kb = bd.A + (size_t)b * ld * ld;")
    w("  const double* al = bd.alpha + (size_t)b * ld;")
    sum_names = [(c, k) for c, cp in enumerate(comps) for k in cp.sums]
    w("  for (;;) {")
    w("    int unit = 0;")
    w("    if (lane == 0) unit = atomicAdd(&sm.next_unit, 1);")
    w("    unit = __shfl_sync(0xffffffffu, unit, 0);")
    w("    if (unit >= n_units) break;")
    w("    const int t = t0 + (unit >> 3), reg = unit & 7;")
    w("    const int wr = (reg >> 1) * 16, wc = (reg & 1) * 32;          // the region inside tile t")
    w("    int ti, tj;")
    w("    wv_tile_from_linear(t, ti, tj);")
    w("    if (ti == tj && wc > wr + 15) {             // strictly above the diagonal: contributes nothing")
    w(f"      for (int k = lane; k < {nsum}; k += 32) sm.wsum[t - t0][reg][k] = 0.0;")
    w("      continue;")
    w("    }")
    stage_warp(True)
    w("    double trw = 0.0;")
    if sum_names:
        w("    double " + ", ".join(f"s{c}_{k} = 0.0" for c, k in sum_names) + ";")
    w("    {")
    w("      double aj[4];")
    w("      wvs_ld4(&sw.al[c_off], aj);")
    w("      const int gj0 = tj * 64 + wc + (c_off - 16);")
    w("      const double* Krow = Kb + (size_t)(ti * 64 + wr + r_loc) * ld + gj0;")
    w("      double2 n01 = *reinterpret_cast<const double2*>(Krow), n23 = *reinterpret_cast<const double2*>(Krow + 2);")
    w("#pragma unroll 1")
    w("      for (int a = 0; a < 4; ++a) {")
    w("        const int r = r_loc + a;")
    w("""        double w[4];
        {
          const double kin[4] = {n01.x, n01.y, n23.x, n23.y};
          if (a < 3) {       // next row's K^-1 entries: in flight while this row is reduced
            n01 = *reinterpret_cast<const double2*>(Krow + (size_t)(a + 1) * ld);
            n23 = *reinterpret_cast<const double2*>(Krow + (size_t)(a + 1) * ld + 2);
          }
          const int gi = ti * 64 + wr + r;
          const double ai = sw.al[r];
#pragma unroll
          for (int bb = 0; bb < 4; ++bb) {
            const int gj = gj0 + bb;
            const double wv = ai * aj[bb] - kin[bb];
            const bool in = gi < n && gj < n;
            w[bb] = in ? (gi > gj ? 2.0 * wv : (gi == gj ? wv : 0.0)) : 0.0;
            if (gi == gj && gi < n) trw += wv;
          }
        }""")
    cat_loads("        ")
    for c, cp in enumerate(comps):
        if not cp.sums:
            continue
        w((f"        if (cmask & {1 << c}u) {{" if USE_CMASK else "        {") +
          f"      // component {c}: sums {', '.join(f'{k} -> {v}' for k, v in cp.sums.items())}")
        ind = "          "
        skip = bool(cp.cats) and bool(cp.se or cp.other)
        if skip:
            for s_ in _warp_skip_open(cp, ind):
                w(s_)
            ind += "  "
        for s_ in _emit_loads(cp, ind):
            w(s_)
        w("#pragma unroll")
        w(f"{ind}for (int j = 0; j < 4; ++j) {{")
        w(f"{ind}  double v = w[j];")
        if cp.cats:
            w(f"{ind}  if (!({_mask_expr(cp, 'j')})) v = 0.0;")
        if cp.se:
            for i, (x, _) in enumerate(cp.se):
                w(f"{ind}  const double d{i} = xi{x} - xj{x}[j], u{i} = d{i} * d{i};")
                w(f"{ind}  {'double un = -u0;' if i == 0 else f'un -= u{i};'}")
            w(f"{ind}  v *= wv_exp2_big_lo(un, sm.tab);")
        for s_ in value_terms(cp, "a", "j", True):
            w(f"{ind}  " + s_)
        if "S" in cp.sums:
            w(f"{ind}  s{c}_S += v;")
        for i in range(len(cp.se)):
            if f"se{i}" in cp.sums:
                w(f"{ind}  s{c}_se{i} = fma(v, u{i}, s{c}_se{i});")
        for i, o in enumerate(cp.other):
            if f"ls{i}" in cp.sums:
                w(f"{ind}  s{c}_ls{i} = fma(v, q{i}, s{c}_ls{i});")
            if f"aux{i}" in cp.sums:
                w(f"{ind}  s{c}_aux{i} = fma(v, qa{i}, s{c}_aux{i});")
        w(f"{ind}}}")
        if skip:
            w("          }")
        w("        }")
    w("      }")
    w("    }")
    w("    __syncwarp();       // every lane is done with the staged columns: `red` aliases them")
    w(f"    sw.red[{trw_sum} * 33 + lane] = trw;")
    for c, cp in enumerate(comps):
        for k, idx in cp.sums.items():
            w(f"    sw.red[{idx} * 33 + lane] = s{c}_{k};")
    w("    __syncwarp();")
    w(f"    for (int k = lane; k < {nsum}; k += 32) sm.wsum[t - t0][reg][k] = wvs_warp_row_sum(sw.red, k);")
    w("  }")
    w("  __syncthreads();")
    w(f"  for (int i = threadIdx.x; i < (t1 - t0) * {nsum}; i += WVS_THREADS) {{")
    w(f"    const int tl = i / {nsum}, k = i % {nsum};")
    w("    double s_ = sm.wsum[tl][0][k];")
    w("#pragma unroll")
    w("    for (int q = 1; q < WVS_THREADS / 32; ++q) s_ += sm.wsum[tl][q][k];")
    w("    sm.sums[tl][k] = s_;")
    w("  }")
    w("  __syncthreads();")
    # ---- scalar epilogue: partial[slot] = sum of (reduced sum) x (scalar coefficient)
    contrib: Dict[int, List[str]] = {s: [] for s in range(ns)}
    for cp in comps:
        if not cp.sums:
            continue
        th = [f"sm.theta[{s}]" for s in cp.var_slots]
        var_all = _prod(th)
        if "S" in cp.sums:
            for k, s in enumerate(cp.var_slots):
                if trainable[s]:
                    contrib[s].append(f"S[{cp.sums['S']}] * ({_prod(th[:k] + th[k + 1:])})")
        for i, (_, sl) in enumerate(cp.se):
            if f"se{i}" in cp.sums:
                contrib[sl].append(f"S[{cp.sums[f'se{i}']}] * ({var_all}) * ({TWO_LN2} / sm.theta[{sl}])")
        for i, o in enumerate(cp.other):
            if f"ls{i}" in cp.sums:
                contrib[o["ls"]].append(f"S[{cp.sums[f'ls{i}']}] * ({var_all}) / sm.theta[{o['ls']}]")
            if f"aux{i}" in cp.sums:
                contrib[o["aux"]].append(f"S[{cp.sums[f'aux{i}']}] * ({var_all}) / (sm.theta[{o['ls']}] * sm.theta[{o['aux']}])")
    contrib[int(p.noise_slot)].append(f"S[{trw_sum}]")
    w(f"  for (int i = threadIdx.x; i < (t1 - t0) * {ns}; i += WVS_THREADS) {{")
    w(f"    const int tl = i / {ns}, slot = i % {ns};")
    w("    const double* S = sm.sums[tl];")
    w("    double v = 0.0;")
    w("    switch (slot) {")
    for s_ in range(ns):
        if contrib[s_]:
            w(f"      case {s_}: v = {' + '.join(contrib[s_])}; break;")
    w("      default: break;")
    w("    }")
    w("    bd.partial[((size_t)b * ntiles + t0 + tl) * bd.n_slots_max + slot] = v;")
    w("  }")
    w("}")

    source = prelude() + "\n" + "\n".join(body) + "\n"
    return SpecSource(source=source, gram_name=gram_name, grad_name=grad_name, gram_smem=gram_smem, grad_smem=grad_smem,
                      n_sums=nsum, key=hashlib.sha1(source.encode()).hexdigest())
