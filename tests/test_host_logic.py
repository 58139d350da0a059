"""Pipe / Fold combinators of the host-side mirror (composable-sdr_b200/blocks.py) against the semantics of
src/ComposableSDR/Types.hs and Trans.hs, using plain-Python pipes (no GPU)."""
import numpy as np

import composable_sdr_b200 as cs


def counting_pipe(log, name, fn):
    def start():
        log.append(("start", name))
        return {"n": 0}

    def process(r, a):
        r["n"] += 1
        return fn(a)

    def done(r):
        log.append(("done", name, r["n"]))
    return cs.Pipe(start, process, done)


def test_compose_order_and_lifecycle():
    log = []
    p1 = counting_pipe(log, "p1", lambda a: a * 2)
    p2 = counting_pipe(log, "p2", lambda a: a + 1)
    process, cleanup = cs.unPipe(p1 * p2)              # p2 first, then p1 (Types.hs:99)
    out = list(process([np.array([1.0]), np.array([2.0])]))
    cleanup()
    assert [float(o[0]) for o in out] == [4.0, 6.0]
    assert log == [("start", "p1"), ("start", "p2"), ("done", "p2", 2), ("done", "p1", 2)]


def test_take_n_arr_trims_the_crossing_chunk():
    chunks = [np.arange(10), np.arange(10, 20), np.arange(20, 30)]
    got = list(cs.takeNArr(15, iter(chunks)))
    assert [len(c) for c in got] == [10, 5]
    assert got[1][-1] == 14


def test_compact_emits_exact_chunks_and_flushes():
    seen = []
    sink = cs.Fold(lambda s, a: seen.append(len(a)) or s, lambda: None, lambda s: None)
    cs.compact(8, sink).run(iter([np.zeros(5), np.zeros(5), np.zeros(9), np.zeros(2)]))
    assert seen == [8, 8, 5]                            # two exact chunks, then the remainder at the end


def test_mux_and_mix():
    log = []
    pipes = [counting_pipe(log, f"c{i}", (lambda k: (lambda a: a * k))(i + 1)) for i in range(3)]
    m = cs.mix * cs.mux(pipes)
    r = m._start()
    out = m._process(r, [np.ones(4), np.ones(4), np.ones(4)])
    m._done(r)
    assert np.array_equal(out, np.full(4, 6.0))         # 1 + 2 + 3, channel order
    assert [e for e in log if e[0] == "start"] == [("start", "c0"), ("start", "c1"), ("start", "c2")]


def test_add_pipe_and_distribute():
    log = []
    folds = [cs.addPipe(counting_pipe(log, f"d{i}", lambda a: a), cs.listSink()) for i in range(2)]
    res = cs.distribute_(folds).run(iter([[np.ones(2), np.zeros(3)], [np.ones(1), np.zeros(1)]]))
    assert [len(r) for r in res] == [3, 4]


def test_reference_side_patch_script(tmp_path):
    """integration/apply_gpu_blocks.py on a miniature stand-in for a checkout: the cabal link line and the two
    wrapper bodies are replaced, everything around them is kept, a missing anchor is refused"""
    import os
    import subprocess
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    script = os.path.join(root, "integration", "apply_gpu_blocks.py")
    co = tmp_path / "co"
    (co / "src" / "ComposableSDR").mkdir(parents=True)
    (co / "composable-sdr.cabal").write_text("library\n  extra-lib-dirs:      /usr/local/lib\n  extra-libraries:     SoapySDR, liquid\n  other: kept\n")
    chs = ('before\nforeign import ccall unsafe "agc_crcf_destroy" c_agc_crcf_destroy\n  :: Agc -> IO ()\n\n'
           'agcExecuteBlock ::\n     sig\nagcExecuteBlock agc px n py = do\n  OLD AGC BODY\n\nagcCreate :: Float -> IO Agc\nmiddle\n'
           'firpfbchCreate :: Int -> IO x\ncreate body\n\nfirpfbchChan ::\n  sig\nfirpfbchChan (fb, nco, nch) a = do\n  OLD PFB BODY\n\n'
           'firpfbchChannelizer ::\n  after\n')
    (co / "src" / "ComposableSDR" / "Liquid.chs").write_text(chs)
    out = subprocess.run([sys.executable, script, str(co), "--lib-dir", "/opt/x"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    cab = (co / "composable-sdr.cabal").read_text()
    assert "SoapySDR, csdr_liquid_compat, csdr_b200, liquid" in cab and "/usr/local/lib, /opt/x" in cab and "other: kept" in cab
    new = (co / "src" / "ComposableSDR" / "Liquid.chs").read_text()
    assert "OLD AGC BODY" not in new and "OLD PFB BODY" not in new
    assert "c_csdr_agc_squelch_execute_block agc px n py" in new and "c_csdr_firpfbch_execute_block fb nco x" in new
    for kept in ("before", "middle", "create body", "  after", "agcCreate :: Float -> IO Agc", "firpfbchChannelizer ::"):
        assert kept in new
    assert out.stdout.startswith("--- a/composable-sdr.cabal")
    # a tree without the anchors is refused and left alone
    (co / "src" / "ComposableSDR" / "Liquid.chs").write_text("nothing to see\n")
    (co / "composable-sdr.cabal").write_text("library\n  extra-lib-dirs:      /usr/local/lib\n  extra-libraries:     SoapySDR, liquid\n")
    bad = subprocess.run([sys.executable, script, str(co)], capture_output=True, text=True)
    assert bad.returncode != 0 and "anchor not found" in bad.stderr
    assert (co / "src" / "ComposableSDR" / "Liquid.chs").read_text() == "nothing to see\n"
