"""The ZDAU ladder step in ecsimd_b200/csrc/zdau_order.inc is a GENERATED permutation of the field
operations of curve_group.h:120-153 (tools/order_search.py picks the order that makes ptxas emit the
fewest instructions).  Values cannot depend on the order as long as it is a valid topological order of
the data flow and every repair-in-place check (fp_quirk_check) precedes the first use of the squares it
may repair: this test pins exactly that, on the CPU."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import order_search as osr  # noqa: E402


def _order_in_file():
    path = os.path.join(ROOT, "ecsimd_b200", "csrc", "zdau_order.inc")
    names = []
    for line in open(path):
        line = line.strip()
        if not line or line.startswith("//") or line.startswith("QuirkAcc f1"):
            continue
        m = re.match(r"(?:const )?fe(?:512)? (\w+) = ", line)
        if m:
            names.append(m.group(1))
            continue
        m = re.match(r"fp_quirk_check<QUIRK>\(md, f(\d),", line)
        assert m, "unexpected statement: %s" % line
        names.append("chk" + m.group(1))
    return names


def test_order_file_is_a_valid_permutation():
    order = _order_in_file()
    assert sorted(order) == sorted(osr.NAMES)          # every operation exactly once
    assert osr.valid(order)                            # data flow + checks before consumers


def test_statements_match_the_search_tool():
    """the text of each statement in the file is the tool's (with one of the two C4 forms)"""
    path = os.path.join(ROOT, "ecsimd_b200", "csrc", "zdau_order.inc")
    body = [l.strip() for l in open(path) if l.strip() and not l.strip().startswith("//")]
    want = set()
    for n, t in osr.STMTS:
        want.add(t.replace("C4EXPR", "fp_shl2_mulonly(C, md)"))
        want.add(t.replace("C4EXPR", "fp_shl<2>(C, md)"))
    # equivalent forms the tool may choose: commuted operands, the other association of a - b - c
    want.update(osr.ALT.values())
    want.update(t for _, t in osr.ALT_PARTNER.values())
    for l in body[1:]:
        assert l in want, l


def test_double_subtractions_are_consistent():
    """a - b - c may be associated either way, but both statements of the pair must agree"""
    path = os.path.join(ROOT, "ecsimd_b200", "csrc", "zdau_order.inc")
    body = {l.strip() for l in open(path)}
    for first, (second, second_text) in osr.ALT_PARTNER.items():
        assert (osr.ALT[first] in body) == (second_text in body), (first, second)
        assert (osr.TEXT[first] in body) == (osr.TEXT[second] in body), (first, second)


def test_dependencies_cover_the_repair_rule():
    # consumers of a group's squares depend on the group's check
    for grp, chk in ((("Cp", "Dp"), "chk1"), (("C", "s4", "s6"), "chk2"), (("D", "Dc"), "chk3")):
        for n in osr.NAMES:
            if n != chk and n not in grp and any(re.search(r"\b%s\b" % g, osr.TEXT[n].split("=", 1)[-1]) for g in grp):
                assert chk in osr.DEPS[n], (n, chk)
