"""CPU: invariants of the per-kernel shared-memory layouts and of the workspace-row exchange lists libb2s builds (host-only entry point
b2s_debug_layouts; no device needed).  The kernels address their workspace only through these tables, so overlapping or misaligned
regions would corrupt the simulation silently: checked here for every packaged task model, both tail tiers and both controller
placements."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.util import ROOT

FIELDS = ("qpos qvel qacc qacc_ws ctrl xpos xquat xmat xipos cdof cdofdot cinert cvel frne ffl M H bias passive qact qsmooth qaccs qcon "
          "gpos gmat spos smat c_pos c_frame c_dist c_fric c_solref c_solimp c_mu c_int J e_D e_R e_aref e_jar e_jv e_force e_floss e_int "
          "Ma grad search Mv scratch scratch_size fused_stride hdr total mc me").split()
LAY = ["FULL", "P0", "TS", "TL", "ROW"]
NPIO, MAXREG = 5, 12
MODELS = ["Lift_Panda", "Lift_Sawyer", "Stack_Panda", "Stack_Sawyer", "Door_Panda", "NutAssemblyRound_Panda", "PickPlace_Panda"]


def _lib():
    so = os.path.join(ROOT, "robosuite_b200", "libb2s.so")
    if not os.path.exists(so):
        pytest.skip("libb2s.so not built")
    L = C.CDLL(so)
    L.b2s_debug_layouts.argtypes = [C.c_int] * 12 + [C.c_void_p] * 3
    return L


def _dims(name):
    from robosuite_b200.mjcf.compiler import load_model

    m = load_model(os.path.join(ROOT, "robosuite_b200", "assets", "models", name + ".npz"))
    pair = np.asarray(m.pair_geom).reshape(-1, 2)
    cg = sorted(set(pair.flatten().tolist()))
    maxdim = int(np.asarray(m.geom_condim)[cg].max())
    return m, dict(nq=m.nq, nv=m.nv, nu=m.nu, nb=m.nbody, ncg=len(cg), ns=m.nsite, hc=maxdim * maxdim)


def layouts(L, d, mc, me, mcs, mes, osc):
    words = (C.c_int * 5)()
    lay = (C.c_int * (5 * len(FIELDS)))()
    pio = (C.c_int * (NPIO * (3 + 2 * MAXREG * 4)))()
    rc = L.b2s_debug_layouts(d["nq"], d["nv"], d["nu"], d["nb"], d["ncg"], d["ns"], d["hc"], mc, me, mcs, mes, int(osc), words, lay, pio)
    assert rc == 0
    lays = {LAY[k]: dict(zip(FIELDS, list(lay)[k * len(FIELDS):(k + 1) * len(FIELDS)])) for k in range(5)}
    pios = []
    for k in range(NPIO):
        o = list(pio)[k * (3 + 2 * MAXREG * 4):(k + 1) * (3 + 2 * MAXREG * 4)]
        ld = [tuple(o[3 + 4 * i:3 + 4 * i + 4]) for i in range(o[0])]
        st = [tuple(o[3 + 4 * MAXREG + 4 * i:3 + 4 * MAXREG + 4 * i + 4]) for i in range(o[1])]
        pios.append(dict(load=ld, store=st, load_words=o[2]))
    return list(words), lays, pios


def sizes(d, mc, me):
    nq, nv, nu, nb, ncg, ns, hc = (d[k] for k in ("nq", "nv", "nu", "nb", "ncg", "ns", "hc"))
    return dict(qpos=nq, qvel=nv, qacc=nv, qacc_ws=nv, ctrl=nu, xpos=3 * nb, xquat=4 * nb, xmat=9 * nb, xipos=3 * nb, cdof=6 * nv,
                cdofdot=6 * nv, cinert=10 * nb, cvel=6 * nb, frne=6 * nb, ffl=6 * nb, M=nv * nv, H=nv * nv, bias=nv, passive=nv, qact=nv,
                qsmooth=nv, qaccs=nv, qcon=nv, gpos=3 * ncg, gmat=9 * ncg, spos=3 * ns, smat=9 * ns, c_pos=3 * mc, c_frame=3 * mc, c_dist=mc,
                c_fric=3 * mc, c_int=5 * mc, J=me * nv, e_D=me, e_R=me, e_aref=me, e_jar=me, e_jv=me, e_force=me, e_floss=me, e_int=me,
                Ma=nv, grad=nv, search=nv, Mv=nv, hdr=8)


def check_disjoint(tag, lay, names, sz, allowed=()):
    spans = []
    for n in names:
        a, b = lay[n], lay[n] + sz[n]
        assert a % 4 == 0, (tag, n, "offset not 16-byte aligned")
        assert 0 <= a and b <= lay["total"], (tag, n, a, b, lay["total"])
        spans.append((a, b, n))
    spans.append((lay["scratch"], lay["scratch"] + lay["scratch_size"], "scratch"))
    assert lay["scratch"] + lay["scratch_size"] <= lay["total"]
    for i in range(len(spans)):
        for j in range(i + 1, len(spans)):
            a, b = spans[i], spans[j]
            if a[0] < b[1] and b[0] < a[1]:
                assert (a[2], b[2]) in allowed or (b[2], a[2]) in allowed, (tag, "overlap", a, b)


P0_LIVE = "qpos qvel xpos xquat xmat xipos cdof cdofdot cinert cvel bias passive gpos gmat spos smat hdr".split()
TAIL_CORE = ("qpos qvel qacc qacc_ws ctrl cdof M H bias passive qact qsmooth qaccs qcon c_pos c_frame c_dist c_fric c_int e_D e_R e_aref e_jar "
             "e_jv e_force e_floss e_int Ma grad search Mv J hdr").split()
ROW_REGS = "xpos xquat cdof cvel M bias passive spos smat gpos gmat hdr".split()


@pytest.mark.parametrize("name", MODELS)
@pytest.mark.parametrize("osc_in_tail", [False, True])
def test_layout_invariants(name, osc_in_tail):
    L = _lib()
    m, d = _dims(name)
    for (mc, me, mcs, mes) in ((32, 64, 8, 32), (96, 288, 32, 96), (32, 64, 32, 64)):
        words, lays, pios = layouts(L, d, mc, me, mcs, mes, osc_in_tail)
        szL, szS = sizes(d, mc, me), sizes(d, mcs, mes)
        # phase 0: M may live over frne + ffl (dead once crb runs); frne / ffl themselves must not collide with anything live
        p0 = lays["P0"]
        check_disjoint(name + "/P0", p0, P0_LIVE + ["frne", "ffl"], szL)
        mlo, mhi = p0["M"], p0["M"] + szL["M"]
        for n in P0_LIVE:
            assert not (p0[n] < mhi and mlo < p0[n] + szL[n]), (name, "P0: M overlaps live region", n)
        assert mhi <= p0["total"] and p0["scratch_size"] >= max(10 * d["nb"], 200)
        # tail tiers
        for tier, sz, cap in (("TS", szS, (mcs, mes)), ("TL", szL, (mc, me))):
            t = lays[tier]
            assert (t["mc"], t["me"]) == cap
            late = ["xpos", "xquat", "spos", "smat"]
            if osc_in_tail:
                check_disjoint(name + "/" + tier, t, TAIL_CORE + late + ["cvel"], sz)
                assert t["scratch_size"] >= 672
            else:
                check_disjoint(name + "/" + tier, t, TAIL_CORE + late, sz, allowed=[(x, "J") for x in late])
            hs = cap[1] + d["hc"] * cap[0] + 64
            assert t["scratch_size"] >= max(hs, 9 * cap[0]), (name, tier, t["scratch_size"], hs)
        # row
        check_disjoint(name + "/ROW", dict(lays["ROW"], scratch=0, scratch_size=0), ROW_REGS, szL)
        # exchange lists: inside both address spaces, 16-byte aligned, byte counts consistent
        srcs = {0: "P0", 1: "TS", 2: "TS", 3: "TL", 4: "TL"}
        for k, io in enumerate(pios):
            tot = lays[srcs[k]]["total"]
            for (off, goff, ln, dyn) in io["load"] + io["store"]:
                assert off % 4 == 0 and goff % 4 == 0 and ln % 4 == 0 and ln > 0 and dyn == 0
                assert off + ln <= tot and goff + ln <= lays["ROW"]["total"], (name, k, off, goff, ln)
            assert io["load_words"] == sum(r[2] for r in io["load"])
        assert len(pios[0]["store"]) >= 1 and len(pios[0]["load"]) == 0
        need = sum((szL[n] + 3) // 4 * 4 for n in ROW_REGS[:-1])
        assert sum(r[2] for r in pios[0]["store"]) == need
        want_early = ["cdof", "cvel", "M", "bias", "passive", "xpos", "xquat", "spos", "smat"] if osc_in_tail else ["cdof", "M", "bias", "passive"]
        for k in (1, 3):
            assert pios[k]["load_words"] == sum((szL[n] + 3) // 4 * 4 for n in want_early)
        for k in (2, 4):
            assert pios[k]["load_words"] == (0 if osc_in_tail else sum((szL[n] + 3) // 4 * 4 for n in ("xpos", "xquat", "spos", "smat")))
        # every exchanged region maps the SAME row words in producer and consumer
        row = lays["ROW"]
        for k, io in enumerate(pios):
            lay = lays[srcs[k]]
            for (off, goff, ln, _) in io["load"] + io["store"]:
                names = [n for n in ROW_REGS[:-1] if row[n] >= goff and row[n] < goff + ln]
                for n in names:
                    assert lay[n] - off == row[n] - goff, (name, k, n)


def test_words_per_warp_report():
    """prints the occupancy every task gets (words per warp and warps per SM at 228 KB, fp32)"""
    L = _lib()
    caps = {"Lift": (32, 64), "Stack": (32, 64), "Door": (48, 160), "NutAssemblyRound": (96, 288), "PickPlace": (64, 224)}
    for name in MODELS:
        m, d = _dims(name)
        mc, me = caps[name.split("_")[0]]
        words, lays, pios = layouts(L, d, mc, me, max(4, mc // 4), max(me // 3, 24), False)
        print("%-24s fused %5d  P0 %5d (%2d warps/SM)  tail small %5d (%2d)  tail large %5d (%2d)  row %5d" % (
            name, words[0], words[1], 228 * 256 // words[1], words[2], 228 * 256 // words[2], words[3], 228 * 256 // words[3], words[4]))
