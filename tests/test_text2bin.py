"""`dw text2bin` / `dw bin2text` against the reference's own fixtures:
  * test/text2bin.bats:11-27 -- dd_{variables,factors,weights}.txt with their xxd dumps *.bin.txt
  * every TSV of the seven bundled graphs (test/<name>/*.tsv + *.text2bin-args) against the bytes
    and the record count the UNMODIFIED reference's text2bin wrote for it
  * the reference's bin2text rendering of the bundled graphs' binary files
Fixtures: tests/golden/text2bin/ (tests/golden/make_golden.py text2bin).  Host-only: no GPU."""
import os
import re

import pytest

import sampler_b200 as sb
from helpers import GOLDEN, GRAPHS, graph_paths

T2B = os.path.join(GOLDEN, "text2bin")
KIND = {"variables": "variable", "domains": "domain", "factors": "factor", "weights": "weight"}


def _xxd_bytes(path):
    """bytes of an `xxd` dump (compare_binary_with_xxd_text, test/helpers.bash)"""
    out = bytearray()
    for line in open(path):
        m = re.match(r"^[0-9a-f]+:\s((?:[0-9a-f]{2,4}\s)+)", line)
        if m:
            out += bytes.fromhex(m.group(1).replace(" ", ""))
    return bytes(out)


@pytest.mark.parametrize("what,extra", [("variable", []), ("factor", ["2", "1", "1"]), ("weight", [])])
def test_reference_bats_fixtures(what, extra, tmp_path):
    src = os.path.join(T2B, "dd_%ss.txt" % what)
    out, cnt = tmp_path / "out.bin", tmp_path / "count"
    assert sb.dw_main(["dw", "text2bin", what, src, str(out), str(cnt)] + extra) == 0
    assert out.read_bytes() == _xxd_bytes(os.path.join(T2B, "dd_%ss.bin.txt" % what))


def _tsvs():
    for g in GRAPHS:
        d = os.path.join(T2B, g)
        for f in sorted(os.listdir(d)):
            if f.endswith(".tsv"):
                yield g, f[:-4]


@pytest.mark.parametrize("graph,stem", list(_tsvs()))
def test_text2bin_bundled_graphs(graph, stem, tmp_path):
    d = os.path.join(T2B, graph)
    what = KIND[re.match(r"[a-z]+", stem).group(0)]
    args = []
    a = os.path.join(d, stem + ".text2bin-args")
    if os.path.exists(a):
        args = open(a).read().split()
    out, cnt = tmp_path / "out.bin", tmp_path / "count"
    assert sb.dw_main(["dw", "text2bin", what, os.path.join(d, stem + ".tsv"), str(out), str(cnt)] + args) == 0
    assert out.read_bytes() == open(os.path.join(d, stem + ".bin"), "rb").read()
    assert cnt.read_text().split() == open(os.path.join(d, stem + ".count")).read().split()


@pytest.mark.parametrize("graph", GRAPHS)
def test_bin2text_bundled_graphs(graph, tmp_path):
    p = graph_paths(graph)
    assert sb.dw_main(["dw", "bin2text", "-m", p["meta"], "-v", p["variables"], "-f", p["factors"], "-w", p["weights"],
                       "--domains", p["domains"], "-o", str(tmp_path)]) == 0
    # (domains.tsv is not compared: the reference's dump_domains, src/bin2text.cc:52-69, prints
    # FactorGraph::values without ever building them -- uninitialised memory, different on every
    # run; ours prints the domain values of the file, checked against the input below)
    # The same call prints the equal_to value of an edge to a categorical variable (dump_factors,
    # :83-90), so factors.tsv is compared on the Boolean-only graphs.
    ref = os.path.join(T2B, graph, "bin2text")
    src = os.path.join(T2B, graph, "domains.tsv")
    categorical = any(l.split("\t")[3].strip() == "1" for l in open(os.path.join(ref, "variables.tsv")) if l.strip())
    for f in sorted(os.listdir(ref)):
        if f == "factors.tsv" and categorical:
            continue
        assert (tmp_path / f).read_text() == open(os.path.join(ref, f)).read(), f
    if os.path.exists(src) and os.path.getsize(src):
        want = [l.split("\t")[:3] for l in open(src).read().splitlines() if l.strip()]
        # one line per categorical variable (src/bin2text.cc:54-56): the ones of the input must be there
        got = {l.split("\t")[0]: l.split("\t")[:3] for l in (tmp_path / "domains.tsv").read_text().splitlines() if l.strip()}
        assert all(got[w[0]] == w for w in want)
