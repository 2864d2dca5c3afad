"""Pure-Python reading of the compiled device image (graph_image.h): evaluates potentials
exactly the way the sweep kernels walk the layout.  Test infrastructure: lets the CPU suite
check compile.cc's layout against the oracle without a GPU."""
import math

import numpy as np

TAG_FUNC_MASK, TAG_SELF_EQ1, TAG_SELF_HEAD, TAG_FORM_SHIFT, FORM_GENERAL = 0x0F, 0x10, 0x20, 6, 3
FC_IMPLY_NATURAL, FC_OR, FC_AND, FC_EQUAL, FC_LINEAR, FC_RATIO, FC_LOGICAL, FC_AND_CAT, FC_IMPLY_MLN = range(9)
VF_EVID, VF_CAT, VF_TRUTH = 1, 2, 4


def sign(fc, n, nsb, head):
    nbody = n - 1
    body_all = nsb == nbody
    if fc == FC_AND: return 1.0 if (body_all and head) else -1.0
    if fc == FC_AND_CAT: return 1.0 if (body_all and head) else 0.0
    if fc == FC_OR: return 1.0 if (nsb > 0 or head) else -1.0
    if fc == FC_EQUAL:
        t = nsb + (1 if head else 0)
        return 1.0 if t in (0, n) else -1.0
    if fc == FC_IMPLY_NATURAL: return (1.0 if head else -1.0) if body_all else 0.0
    if fc == FC_IMPLY_MLN: return (1.0 if head else 0.0) if body_all else 1.0
    if n == 1: return 1.0 if head else 0.0
    res = float(nbody) if head else float(nbody - nsb)
    if fc == FC_LINEAR: return res
    if fc == FC_RATIO: return math.log2(1.0 + res)
    return 1.0 if res > 0 else 0.0


class ImageModel:
    def __init__(self, graph):
        g = graph
        self.info = g.info
        for name in ("perm", "inv", "color_start", "var_flags", "evid_val", "var_row", "row_ptr", "row_oth",
                     "truthiness", "e_wid", "e_fv32", "e_fv64", "e_tag", "oth", "row_of_value", "values",
                     "value_base", "w_perm", "w_inv", "sl_off", "blob"):
            setattr(self, name, g.image(name))
        self.packed = bool(g.image("oth_packed")[0])
        self.nvar = int(g.info.num_variables)
        self.nvar_pad = len(self.perm)
        self.valid = self.perm != 0xFFFFFFFF

    def to_schedule_order(self, a_orig):
        a = np.zeros(self.nvar_pad, np.uint64)
        a[self.inv.astype(np.int64)] = np.asarray(a_orig, np.uint64)
        return a

    def row0(self, d):
        return int(self.var_row[d]) if len(self.var_row) else d

    def nrows(self, d):
        return int(self.var_row[d + 1] - self.var_row[d]) if len(self.var_row) else 1

    def fv(self, e):
        if len(self.e_fv32): return float(self.e_fv32[e])
        if len(self.e_fv64): return float(self.e_fv64[e])
        return 1.0

    def entry(self, i):
        if self.packed:
            w = int(self.oth[i]); return w >> 1, w & 1
        return int(self.oth[2 * i]), int(self.oth[2 * i + 1])

    def edge_sign(self, tag, oc, d, cat, key, a, x):
        """-> (sign, span); a = assignments in SCHEDULE order"""
        fc, form = tag & TAG_FUNC_MASK, tag >> TAG_FORM_SHIFT
        if form != FORM_GENERAL:
            self_eq = key if cat else (1 if tag & TAG_SELF_EQ1 else 0)
            self_head = bool(tag & TAG_SELF_HEAD)
            nsb, head_oth = 0, False
            for j in range(form):
                od, eq = self.entry(oc + j)
                s = int(a[od]) == eq
                if (not self_head) and j == form - 1: head_oth = s
                else: nsb += s
            m = x == self_eq
            if self_head: return sign(fc, form + 1, nsb, m), form
            return sign(fc, form + 1, nsb + m, head_oth), form
        n = int(self.oth[oc]) if self.packed else int(self.oth[2 * oc])
        nsb, head = 0, False
        for i in range(n):
            od, eq = self.entry(oc + 1 + i)
            s = (eq == x) if od == d else (int(a[od]) == eq)
            if i == n - 1: head = s
            else: nsb += s
        return sign(fc, n, nsb, head), 1 + n

    def potential(self, vid, proposal, a_orig, weights):
        """FactorGraph::potential(var, proposal) from the image; a_orig in ORIGINAL id order"""
        a = self.to_schedule_order(a_orig)
        weights = np.asarray(weights)[self.w_perm.astype(np.int64)]  # device weight ids are internal
        d = int(self.inv[vid])
        cat = bool(self.var_flags[d] & VF_CAT)
        row = self.row0(d) + (proposal if cat else 0)
        oc = int(self.row_oth[row]) if len(self.row_oth) else 0
        pot = 0.0
        for e in range(int(self.row_ptr[row]), int(self.row_ptr[row + 1])):
            tag = int(self.e_tag[e])
            s, span = self.edge_sign(tag, oc, d, cat, proposal, a, proposal)
            oc += span
            pot += float(weights[self.e_wid[e]]) * (s * self.fv(e))
        return pot


    # ---- the sliced-ELL blob (Boolean graphs): same potentials, read the way k_sell_* do ----
    def slice_header(self, s):
        b = int(self.sl_off[s]) * 8
        h = self.blob[b:b + 8]
        fvw = 1 if self.info.fv_mode == 1 else 2 if self.info.fv_mode == 2 else 0
        n_un, n_mu, n_oc = int(h[0]), int(h[1]), int(h[2])
        un = b + 8
        mu = un + 32 * n_un * (1 + fvw)
        oth = mu + 32 * n_mu * (1 + fvw)
        assert oth + 32 * n_oc == int(self.sl_off[s + 1]) * 8
        return dict(n_un=n_un, n_mu=n_mu, n_oc=n_oc, evid=int(h[3]), pad=int(h[4]), evval=int(h[5]),
                    flags=int(h[6]), un=un, mu=mu, oth=oth, fvw=fvw)

    def _sell_fv(self, sec, ncol, j, lane, fvw):
        if fvw == 0:
            return 1.0
        at = sec + 32 * ncol
        if fvw == 1:
            return float(self.blob[at + 32 * j + lane:at + 32 * j + lane + 1].view(np.float32)[0])
        k = at + 2 * (32 * j + lane)
        return float(self.blob[k:k + 2].view(np.float64)[0])

    def sell_potentials(self, vid, a_orig, weights):
        """(pos, neg) of a Boolean variable from the slice blob: unary records first, then the
        multi-variable ones (sign tables in binary slices, CSR tags + cursor otherwise)"""
        a = self.to_schedule_order(a_orig)
        weights = np.asarray(weights)[self.w_perm.astype(np.int64)]
        d = int(self.inv[vid])
        s, lane = d // 32, d % 32
        h = self.slice_header(s)
        assert not (h["pad"] >> lane) & 1
        assert ((h["evid"] >> lane) & 1) == (1 if self.var_flags[d] & VF_EVID else 0)
        assert ((h["evval"] >> lane) & 1) == int(self.evid_val[d] == 1)
        if h["flags"] & 4:  # SF_WIDE: one variable (slot 0), its records dealt to the 32 lanes: sum the shares
            assert lane == 0 and h["pad"] == 0xFFFFFFFE
            nparts = (h["flags"] >> 18) if h["flags"] & 8 else 1  # SF_PART: a hub cut into consecutive slices
            assert not (h["flags"] & 8) or ((h["flags"] >> 4) & 0x3FFF) == 0
            shares = []
            for q in range(nparts):
                hq = self.slice_header(s + q)
                assert hq["flags"] & 4 and hq["pad"] == 0xFFFFFFFE and hq["evid"] == h["evid"] and hq["evval"] == h["evval"]
                if nparts > 1:
                    assert hq["flags"] & 8 and ((hq["flags"] >> 4) & 0x3FFF) == q and (hq["flags"] >> 18) == nparts
                shares += [self._sell_lane_potentials(hq, l, d, a, weights) for l in range(32)]
            return sum(p for p, _ in shares), sum(n for _, n in shares)
        return self._sell_lane_potentials(h, lane, d, a, weights)

    def _sell_lane_potentials(self, h, lane, d, a, weights):
        pos, neg = 0.0, 0.0
        code = lambda c: float(c - 1)
        antisym = True
        for j in range(h["n_un"]):
            rec = int(self.blob[h["un"] + 32 * j + lane])
            wid = (rec >> 3) & 0x7FFFFF  # unary record: code(1) | fixed << 2 | wid << 3 | code(0) << 26
            w = float(weights[wid])
            fv = self._sell_fv(h["un"], h["n_un"], j, lane, h["fvw"])
            s1, s0 = code(rec & 3), code((rec >> 26) & 3)
            antisym &= (s1 == -s0)
            pos += w * (s1 * fv)
            neg += w * (s0 * fv)
        if h["flags"] & 1:
            assert antisym
        binary = bool(h["flags"] & 2)
        oc = 0
        words = lambda c: int(self.blob[h["oth"] + 32 * c + lane])
        for j in range(h["n_mu"]):
            rec = int(self.blob[h["mu"] + 32 * j + lane])
            tag, wid = rec >> 24, rec & 0x7FFFFF
            w = float(weights[wid])
            fv = self._sell_fv(h["mu"], h["n_mu"], j, lane, h["fvw"])
            if binary:
                wd = words(j)
                sat = int(a[wd >> 1]) == (wd & 1)
                s0 = code((tag >> (2 * sat)) & 3)
                s1 = code((tag >> (4 + 2 * sat)) & 3)
            else:
                form = tag >> TAG_FORM_SHIFT
                if form == 0:
                    continue
                fc = tag & TAG_FUNC_MASK
                if form != FORM_GENERAL:
                    self_eq = 1 if tag & TAG_SELF_EQ1 else 0
                    self_head = bool(tag & TAG_SELF_HEAD)
                    nsb, head_oth = 0, False
                    for k in range(form):
                        wd = words(oc + k)
                        sat = int(a[wd >> 1]) == (wd & 1)
                        if (not self_head) and k == form - 1: head_oth = sat
                        else: nsb += sat
                    oc += form
                    if self_head:
                        s0, s1 = sign(fc, form + 1, nsb, self_eq == 0), sign(fc, form + 1, nsb, self_eq == 1)
                    else:
                        s0, s1 = sign(fc, form + 1, nsb + (self_eq == 0), head_oth), sign(fc, form + 1, nsb + (self_eq == 1), head_oth)
                else:
                    n = words(oc)
                    nsb0 = nsb1 = 0; h0 = h1 = False
                    for k in range(n):
                        wd = words(oc + 1 + k)
                        od, eq = wd >> 1, wd & 1
                        if od == d: t0, t1 = eq == 0, eq == 1
                        else: t0 = t1 = int(a[od]) == eq
                        if k == n - 1: h0, h1 = t0, t1
                        else: nsb0 += t0; nsb1 += t1
                    oc += 1 + n
                    s0, s1 = sign(fc, n, nsb0, h0), sign(fc, n, nsb1, h1)
            pos += w * (s1 * fv)
            neg += w * (s0 * fv)
        return pos, neg

    def sell_fixed_bits_ok(self, w_fixed_orig):
        """every real record's fixed bit == isfixed of its weight (padding: weight id 0, bit clear)"""
        fixed = np.asarray(w_fixed_orig)[self.w_perm.astype(np.int64)].astype(np.uint32)
        for s in range(len(self.sl_off) - 1):
            h = self.slice_header(s)
            recs = self.blob[h["un"]:h["un"] + 32 * h["n_un"]]
            real = recs != 0x04000001  # SELL_UNARY_NOP
            bits = (recs >> 2) & 1
            if not np.array_equal(bits[real], fixed[((recs >> 3) & 0x7FFFFF).astype(np.int64)][real]) or bits[~real].any():
                return False
            nop = 0x55 if h["flags"] & 2 else 0
            recs = self.blob[h["mu"]:h["mu"] + 32 * h["n_mu"]]
            real = (recs >> 24) != nop if nop else ((recs >> 30) != 0)
            bits = (recs >> 23) & 1
            if not np.array_equal(bits[real], fixed[(recs & 0x7FFFFF).astype(np.int64)][real]) or bits[~real].any():
                return False
        return True
