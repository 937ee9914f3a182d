"""Generator check of the warp-cooperative kernel form (PC_GEN_COOP, an opt-in experiment of
pychem_b200/codegen/gen_eri.py, profiles/r2e_cooperative_high_l.txt): the roles must partition the
bra components of the contracted (e0|f0) and the ket functions of the sub-blocks exactly once, and
the generated text must carry one kernel + launcher branch per supported mode.  Numerics of the
form are checked on the host emulation when a variant is built (tools/build_variant.py --emu)."""
import importlib
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture()
def gen(monkeypatch):
    monkeypatch.setenv("PC_GEN_COOP", "dppp=3,dpdp=3,dpds=3")
    sys.path.insert(0, os.path.join(HERE, "..", "pychem_b200", "codegen"))
    import gen_eri
    mod = importlib.reload(gen_eri)
    yield mod
    monkeypatch.delenv("PC_GEN_COOP")
    importlib.reload(gen_eri)
    sys.path.pop(0)


@pytest.mark.parametrize("cls,split", [((2, 1, 1, 1), "d"), ((2, 1, 2, 1), "d"), ((2, 1, 2, 0), "c")])
def test_roles_partition_components_and_sub_blocks(gen, cls, split):
    base = gen.ClassGen(*cls)
    coop = gen.CoopGen(base, 3)
    assert coop.G == 3
    flat = [e for grp in coop.egroups for e in grp]
    assert sorted(flat) == sorted(base.e_list) and len(set(flat)) == len(flat)
    NA, NB, NC, ND = base.nsph
    seen = set()
    for c0, ncs, d0, nds in coop.sub_blocks:
        for c in range(c0, c0 + ncs):
            for d in range(d0, d0 + nds):
                assert (c, d) not in seen
                seen.add((c, d))
    assert len(seen) == NC * ND
    assert coop.defer == (split == "d")
    # every role's recursion is a pruned copy of the class's: no more temporaries than the whole
    base.gen_vrr()
    for grp in coop.egroups:
        sub = coop.sub(grp)
        sub.gen_vrr()
        assert 0 < sub.n_vrr < base.n_vrr


def test_generated_text(gen):
    g = gen.make_class((2, 1, 1, 1))
    src = g.source()
    assert "eri_dppp_coop_kernel" in src and "pc_coop_reduce<MODE, 5, 3, 3, 3>" in src
    for mode in gen.CoopGen.MODES:
        assert "eri_dppp_coop_kernel<%s><<<" % mode in src
    assert "eri_dppp_coop_kernel<PC_MODE_TENSOR>" not in src         # other modes stay on the one-thread kernel
    assert "PYCHEM_B200_COOP" in src
    plain = gen.make_class((1, 1, 1, 0)).source()                    # a class outside PC_GEN_COOP is untouched
    assert "coop" not in plain


def test_multi_pass_classes_partition_the_bra_components():
    """ClassGenPass (the default form of six high-L classes): every bra component of the contracted
    (e0|f0) belongs to exactly one pass, the kernel text has one pair of primitive loops per pass,
    and the flop model is that of the one-pass class (the algorithmic count does not see passes)."""
    sys.path.insert(0, os.path.join(HERE, "..", "pychem_b200", "codegen"))
    try:
        import gen_eri
        gen_eri = importlib.reload(gen_eri)
        assert set(gen_eri.PASS_CLASSES) == {"dppp", "dpdp", "dpds", "ddds", "ddpp", "dddp"}
        for name, npass in gen_eri.PASS_CLASSES.items():
            cls = tuple("spd".index(c) for c in name)
            g = gen_eri.make_class(cls)
            assert isinstance(g, gen_eri.ClassGenPass) and g.groups.G == npass
            flat = [e for grp in g.groups.egroups for e in grp]
            assert sorted(flat) == sorted(g.e_list) and len(set(flat)) == len(flat)
            src = g.source()
            assert src.count("for (int ik = 0; ik < KK;") == npass * len(gen_eri.MODES) // len(gen_eri.MODES)
            plain = gen_eri.ClassGen(*cls)
            plain.source_single()
            assert gen_eri.flop_model(g) == gen_eri.flop_model(plain)
        # the Cartesian-d variant of a class takes the same form (six functions per d shell)
        gc = gen_eri.make_class((2, 1, 1, 1), cart_d=True)
        assert isinstance(gc, gen_eri.ClassGenPass) and gc.name == "Dppp" and gc.nsph == [6, 3, 3, 3]
        assert "g[%d]" % (6 * 3 * 3 * 3 - 1) in gc.source()
        # grouping by an accumulator cap: every pass within the cap
        g = gen_eri.ClassGenPass(2, 2, 2, 1, npass="c112")
        assert all(len(es) * g.nf <= 112 for es in g.groups.egroups) and g.groups.G == 5
    finally:
        sys.path.pop(0)
