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


def _mask_expr(cp: _Comp, a: str, b: str) -> str:
    return " && ".join(f"kr{k}[{a}] == kc{k}[{b}]" for k in cp.cats)


def _emit_loads(cp: _Comp, need_q: bool) -> List[str]:
    out = []
    for k in cp.cats:
        out.append(f"      int kr{k}[4], kc{k}[4]; wvs_ld4i(&sm.cr[{k}][r_off], kr{k}); wvs_ld4i(&sm.cc[{k}][c_off], kc{k});")
    ids = [a for a, _ in cp.se] + [o["arr"] for o in cp.other]
    for a in dict.fromkeys(ids):
        out.append(f"      double xi{a}[4], xj{a}[4]; wvs_ld4(&sm.ar[{a}][r_off], xi{a}); wvs_ld4(&sm.ac[{a}][c_off], xj{a});")
    return out


def _warp_skip_open(cp: _Comp) -> List[str]:
    """whole warps skip the transcendental factors of a product whose categorical mask is zero for the warp"""
    return ["      bool any_ = false;",
            "#pragma unroll",
            "      for (int a = 0; a < 4; ++a)",
            "#pragma unroll",
            f"        for (int b = 0; b < 4; ++b) any_ |= ({_mask_expr(cp, 'a', 'b')});",
            "      if (__any_sync(0xffffffffu, any_)) {"]


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

    def smem_struct(name, with_red):
        w(f"struct {name} {{")
        w(f"  double theta[{ns + (ns & 1)}];")
        w(f"  double kc[{nkc + (nkc & 1)}];")
        w("  double tab[WV_EXP2_TAB12];")
        w(f"  double ar[{na}][64];")
        w(f"  double ac[{na}][64];")
        w(f"  int cr[{ncat}][64];")
        w(f"  int cc[{ncat}][64];")
        if with_red:
            w(f"  double red[{nsum}][WVS_THREADS];")
            w(f"  double sums[{nsum + (nsum & 1)}];")
        w("};")
        size = 8 * (ns + (ns & 1)) + 8 * (nkc + (nkc & 1)) + 8 * 4096 + 2 * 8 * 64 * na + 2 * 4 * 64 * ncat
        if with_red:
            size += 8 * 256 * nsum + 8 * (nsum + (nsum & 1))
        w(f"static_assert(sizeof({name}) == {size}, \"shared-memory layout\");")
        return size

    def stage_model():
        w("  const int b = active[blockIdx.y];")
        w(f"  wvs_stage_theta(bd, b, xall, sm.theta, {ns});")
        w("  wvs_load_tab(sm.tab, gtab);")
        w("  const unsigned cmask = bd.comp_mask[b];")
        w("  __syncthreads();")
        for i, e in enumerate(kc_expr):
            w(f"  if (threadIdx.x == {i % 256}) sm.kc[{i}] = {e};")
        w("  const int n = bd.n, ld = bd.npad;")
        w("  const int t1 = min(ntiles, (int)(blockIdx.x + 1) * tpc);")

    def stage_tile():
        w("    int ti, tj;")
        w("    wv_tile_from_linear(t, ti, tj);")
        w("    __syncthreads();      // previous tile done with the staged columns; sm.kc visible")
        w("    if (threadIdx.x < 128) {")
        w("      const int r = threadIdx.x & 63, col = threadIdx.x >> 6;")
        w("      const size_t g = (size_t)(col ? tj : ti) * 64 + r;")
        dims = sorted({d for (_, d, _) in arrays} | set(cats))
        for d in dims:
            w(f"      const double x{d} = bd.Xt[(size_t){d} * ld + g];")
        for (kind, dim, slot), a in arrays.items():
            sc = f" * sm.kc[{arr_scale[a]}]" if a in arr_scale else ""
            w(f"      (col ? sm.ac : sm.ar)[{a}][r] = x{dim}{sc};")
        for dim, k in cats.items():
            w(f"      (col ? sm.cc : sm.cr)[{k}][r] = (int)rint(x{dim});")
        w("    }")
        w("    __syncthreads();")
        w("    int r_off, c_off;")
        w("    bool above;")
        w("    wvs_coords(r_off, c_off, above, ti == tj);")

    def value_terms(cp: _Comp, a: str, b: str, grads: bool) -> List[str]:
        """statements that multiply the unit-variance values of the non-SE leaves into `v` (and set their q's)"""
        out = []
        for i, o in enumerate(cp.other):
            x = o["arr"]
            if o["type"] == LINEAR:
                out.append(f"v *= xi{x}[{a}] * xj{x}[{b}];")
            elif o["type"] in (M12, M32, M52):
                if grads and f"ls{i}" in cp.sums:
                    out.append(f"double E{i}, q{i}; wvs_matern<{o['type']}>(xi{x}[{a}], xj{x}[{b}], E{i}, q{i}); v *= E{i};")
                else:
                    out.append(f"v *= wvs_matern_value<{o['type']}>(xi{x}[{a}], xj{x}[{b}]);")
            elif o["type"] == PERIODIC:
                if grads and (f"ls{i}" in cp.sums or f"aux{i}" in cp.sums):
                    out.append(f"double E{i}, q{i}, qa{i}; wvs_periodic(xi{x}[{a}], xj{x}[{b}], sm.theta[{o['ls']}], "
                               f"sm.theta[{o['aux']}], E{i}, q{i}, qa{i}); v *= E{i};")
                else:
                    out.append(f"v *= wvs_periodic_value(xi{x}[{a}], xj{x}[{b}], sm.theta[{o['ls']}], sm.theta[{o['aux']}]);")
        return out

    # =============================================================================== gram
    w("// ---------------- generated: Gram builder")
    gram_smem = smem_struct("WvsGramSmem", False)
    w(f"extern \"C\" __global__ void __launch_bounds__(WVS_THREADS, 2) {gram_name}(WvBatchDev bd, const int* __restrict__ active,")
    w("    const double* __restrict__ xall, const double* __restrict__ gtab, int ntiles, int tpc) {")
    w("  WvsGramSmem& sm = *reinterpret_cast<WvsGramSmem*>(wvs_smem_raw);")
    stage_model()
    w("  double* Ab = bd.A + (size_t)b * ld * ld;")
    w("  const double* yb = bd.Y + (size_t)b * ld;")
    w("  const double* lam = bd.site_lam ? bd.site_lam + (size_t)b * ld : nullptr;")
    w("  const double* eta = bd.site_eta ? bd.site_eta + (size_t)b * ld : nullptr;")
    w("  for (int t = blockIdx.x * tpc; t < t1; ++t) {")
    stage_tile()
    w("    if (above) continue;                  // the strict upper part of a diagonal tile is never read")
    w("    double acc[WVS_NE];")
    w("#pragma unroll")
    w("    for (int e = 0; e < WVS_NE; ++e) acc[e] = 0.0;")
    for c, cp in enumerate(comps):
        expensive = bool(cp.se or cp.other)
        w(f"    if (cmask & {1 << c}u) {{      // component {c}")
        for s in _emit_loads(cp, False):
            w(s)
        skip = bool(cp.cats) and expensive
        if skip:
            for s in _warp_skip_open(cp):
                w(s)
        w(f"      const double k_ = sm.kc[{comp_kc[c]}];")
        w("#pragma unroll")
        w("      for (int a = 0; a < 4; ++a)")
        w("#pragma unroll")
        w("        for (int b = 0; b < 4; ++b) {")
        if cp.se:
            w("          double u = k_;")
            for (x, _) in cp.se:
                w(f"          {{ const double d = xi{x}[a] - xj{x}[b]; u = fma(-d, d, u); }}")
            w("          double v = wv_exp2_12_lo(u, sm.tab);")
        else:
            w("          double v = k_;")
        for s in value_terms(cp, "a", "b", False):
            w("          " + s)
        if cp.cats:
            w(f"          if ({_mask_expr(cp, 'a', 'b')}) acc[a * 4 + b] += v;")
        else:
            w("          acc[a * 4 + b] += v;")
        w("        }")
        if skip:
            w("      }")
        w("    }")
    w("    const double s2 = sm.theta[%d];" % int(p.noise_slot))
    w("    const double cmean = %s;" % (f"sm.theta[{int(p.mean_slot)}]" if int(p.mean_slot) >= 0 else "0.0"))
    w("""#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int gi = ti * 64 + r_off + a;
      double out[4];
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int gj = tj * 64 + c_off + bb;
        double v;
        if (gi < n && gj < n) v = acc[a * 4 + bb] + (gi == gj ? (lam ? bd.jitter + 1.0 / lam[gi] : s2) : 0.0);
        else if (gi == n && gj < n) v = (lam ? eta[gj] / lam[gj] : yb[gj]) - cmean;     // RHS row d^T
        else v = (gi == gj) ? 1.0 : 0.0;                     // identity padding (incl. A[n][n] = 1)
        out[bb] = v;
      }
      double2* dst = reinterpret_cast<double2*>(Ab + (size_t)gi * ld + tj * 64 + c_off);
      dst[0] = make_double2(out[0], out[1]);
      dst[1] = make_double2(out[2], out[3]);
    }
  }
}
""")

    # =============================================================================== grad
    w("// ---------------- generated: gradient reduction")
    grad_smem = smem_struct("WvsGradSmem", True)
    w(f"extern \"C\" __global__ void __launch_bounds__(WVS_THREADS, 2) {grad_name}(WvBatchDev bd, const int* __restrict__ active,")
    w("    const double* __restrict__ xall, const double* __restrict__ gtab, int ntiles, int tpc) {")
    w("  WvsGradSmem& sm = *reinterpret_cast<WvsGradSmem*>(wvs_smem_raw);")
    stage_model()
    w("  const double* Kb = bd.A + (size_t)b * ld * ld;")
    w("  const double* al = bd.alpha + (size_t)b * ld;")
    w("  for (int t = blockIdx.x * tpc; t < t1; ++t) {")
    stage_tile()
    w("""    double w[WVS_NE];
    double trw = 0.0;
    if (!above) {
      double aj[4];
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) aj[bb] = al[tj * 64 + c_off + bb];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int gi = ti * 64 + r_off + a;
        const double2* src = reinterpret_cast<const double2*>(Kb + (size_t)gi * ld + tj * 64 + c_off);
        const double2 k01 = src[0], k23 = src[1];
        const double kin[4] = {k01.x, k01.y, k23.x, k23.y};
        const double ai = al[gi];
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          const int gj = tj * 64 + c_off + bb;
          const double wv = ai * aj[bb] - kin[bb];
          const bool in = gi < n && gj < n;
          w[a * 4 + bb] = in ? (gi > gj ? 2.0 * wv : (gi == gj ? wv : 0.0)) : 0.0;
          if (gi == gj && gi < n) trw += wv;
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < WVS_NE; ++e) w[e] = 0.0;
    }""")
    w(f"    sm.red[{trw_sum}][threadIdx.x] = trw;")
    for c, cp in enumerate(comps):
        if not cp.sums:
            continue
        names = list(cp.sums)
        w(f"    {{      // component {c}: sums {', '.join(f'{k} -> {v}' for k, v in cp.sums.items())}")
        w("      double " + ", ".join(f"s_{k} = 0.0" for k in names) + ";")
        w(f"      if (!above && (cmask & {1 << c}u)) {{")
        for s in _emit_loads(cp, True):
            w("  " + s)
        skip = bool(cp.cats)
        if skip:
            for s in _warp_skip_open(cp):
                w("  " + s)
        w("#pragma unroll")
        w("        for (int a = 0; a < 4; ++a)")
        w("#pragma unroll")
        w("          for (int b = 0; b < 4; ++b) {")
        w("            double v = w[a * 4 + b];")
        if cp.cats:
            w(f"            if (!({_mask_expr(cp, 'a', 'b')})) v = 0.0;")
        if cp.se:
            first = True
            for i, (x, _) in enumerate(cp.se):
                w(f"            const double d{i} = xi{x}[a] - xj{x}[b], u{i} = d{i} * d{i};")
                w(f"            {'double un = -u0;' if first else f'un -= u{i};'}")
                first = False
            w("            v *= wv_exp2_12_lo(un, sm.tab);")
        for s in value_terms(cp, "a", "b", True):
            w("            " + s)
        if "S" in cp.sums:
            w("            s_S += v;")
        for i in range(len(cp.se)):
            if f"se{i}" in cp.sums:
                w(f"            s_se{i} = fma(v, u{i}, s_se{i});")
        for i, o in enumerate(cp.other):
            if f"ls{i}" in cp.sums:
                w(f"            s_ls{i} = fma(v, q{i}, s_ls{i});")
            if f"aux{i}" in cp.sums:
                w(f"            s_aux{i} = fma(v, qa{i}, s_aux{i});")
        w("          }")
        if skip:
            w("        }")
        w("      }")
        for k in names:
            w(f"      sm.red[{cp.sums[k]}][threadIdx.x] = s_{k};")
        w("    }")
    w("    __syncthreads();")
    w(f"    wvs_reduce_sums(&sm.red[0][0], {nsum}, sm.sums);")
    w("    __syncthreads();")
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
                    contrib[s].append(f"sm.sums[{cp.sums['S']}] * ({_prod(th[:k] + th[k + 1:])})")
        for i, (_, sl) in enumerate(cp.se):
            if f"se{i}" in cp.sums:
                contrib[sl].append(f"sm.sums[{cp.sums[f'se{i}']}] * ({var_all}) * ({TWO_LN2} / sm.theta[{sl}])")
        for i, o in enumerate(cp.other):
            if f"ls{i}" in cp.sums:
                contrib[o["ls"]].append(f"sm.sums[{cp.sums[f'ls{i}']}] * ({var_all}) / sm.theta[{o['ls']}]")
            if f"aux{i}" in cp.sums:
                contrib[o["aux"]].append(f"sm.sums[{cp.sums[f'aux{i}']}] * ({var_all}) / (sm.theta[{o['ls']}] * sm.theta[{o['aux']}])")
    contrib[int(p.noise_slot)].append(f"sm.sums[{trw_sum}]")
    w(f"    if (threadIdx.x < {ns}) {{")
    w("      double v = 0.0;")
    w("      switch (threadIdx.x) {")
    for s in range(ns):
        if contrib[s]:
            w(f"        case {s}: v = {' + '.join(contrib[s])}; break;")
    w("        default: break;")
    w("      }")
    w("      bd.partial[((size_t)b * ntiles + t) * bd.n_slots_max + threadIdx.x] = v;")
    w("    }")
    w("  }")
    w("}")

    source = prelude() + "\n" + "\n".join(body) + "\n"
    return SpecSource(source=source, gram_name=gram_name, grad_name=grad_name, gram_smem=gram_smem, grad_smem=grad_smem,
                      n_sums=nsum, key=hashlib.sha1(source.encode()).hexdigest())
