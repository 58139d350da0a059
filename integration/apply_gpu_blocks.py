#!/usr/bin/env python3
"""Reference-side change for the GPU blocks: edits a checkout of mryndzionek/composable-sdr in place.

  python integration/apply_gpu_blocks.py /path/to/composable-sdr [--lib-dir /opt/composable-sdr_b200] [--dry-run]

1. composable-sdr.cabal: links libcsdr_liquid_compat + libcsdr_b200 in front of libliquid (INTEGRATION.md option A), so every
   liquid symbol of the hot path resolves to the GPU library and everything else still comes from liquid.
2. src/ComposableSDR/Liquid.chs (INTEGRATION.md option B): two foreign imports, and the two wrappers that call liquid once
   per sample / per frame become one coarse call each:
     agcExecuteBlock  (execute + squelch_get_status + poke 0 per sample)  ->  csdr_agc_squelch_execute_block
     firpfbchChan     (mix_block_down + analyzer_execute per frame + per-element pokes)  ->  csdr_firpfbch_execute_block
   Types, Pipe construction and every caller stay as they are.
The script locates the definitions by name (it carries none of the reference's code), refuses to run when an anchor is
missing, and prints a unified diff of what it changed.  It was exercised against the reference revision this repository
was built from; the result could not be compiled here (no GHC in the build image)."""
import argparse
import difflib
import os
import re
import sys

AGC_IMPORT = """
foreign import ccall unsafe "csdr_agc_squelch_execute_block" c_csdr_agc_squelch_execute_block
  :: Agc -> Ptr SamplesIQCF32 -> CUInt -> Ptr SamplesIQCF32 -> IO CInt
"""

AGC_BODY = """agcExecuteBlock agc px n py =
  -- GPU: gain loop, squelch state machine and gate (status /= SIGNALHI => 0) in one call
  void $ c_csdr_agc_squelch_execute_block agc px n py

"""

PFB_IMPORT = """foreign import ccall unsafe "csdr_firpfbch_execute_block" c_csdr_firpfbch_execute_block
  :: FirPfbch -> Nco -> Ptr SamplesIQCF32 -> CUInt -> Ptr SamplesIQCF32 -> IO CInt

"""

PFB_BODY = """firpfbchChan (fb, nco, nch) a = do
  let nx = A.length a
      nf = nx `div` nch
      x = castPtr . unsafeForeignPtrToPtr $ AT.aStart a
  fy <- mallocPlainForeignPtrBytes (8 * nx)
  withForeignPtr fy $ \\y -> do
    -- GPU: pre-rotation, all nf frames and the channel-major transposition in one call
    _ <- c_csdr_firpfbch_execute_block fb nco x (fromIntegral nx) y
    let v =
          AT.Array
            { AT.aStart = fy
            , AT.aEnd = y `plusPtr` (8 * nf * nch)
            , AT.aBound = y `plusPtr` (8 * nx)
            }
        go as = do
          as' <- as
          if A.length as' == nf
            then return (as', Nothing)
            else let (as'', bs) = AT.splitAt nf as'
                  in return (as'', Just bs)
    return $ unfoldr go (Just v)

"""


def need(cond, what):
    if not cond:
        sys.exit(f"apply_gpu_blocks: anchor not found: {what} (different revision of the reference?)")


def replace_between(text, start_pat, end_pat, new, what):
    """replace text from the line matching start_pat up to (not including) the line matching end_pat"""
    a = re.search(start_pat, text, re.M)
    need(a, what + " (start)")
    b = re.search(end_pat, text[a.start():], re.M)
    need(b, what + " (end)")
    return text[:a.start()] + new + text[a.start() + b.start():]


def insert_before(text, pat, new, what):
    a = re.search(pat, text, re.M)
    need(a, what)
    return text[:a.start()] + new + text[a.start():]


def insert_after(text, pat, new, what):
    a = re.search(pat, text, re.M)
    need(a, what)
    return text[:a.end()] + new + text[a.end():]


def patch_cabal(text, lib_dir):
    need(re.search(r"^\s*extra-libraries:\s*SoapySDR,\s*liquid\s*$", text, re.M), "cabal extra-libraries")
    text = re.sub(r"^(\s*extra-libraries:\s*)SoapySDR,\s*liquid\s*$", r"\1SoapySDR, csdr_liquid_compat, csdr_b200, liquid", text, count=1, flags=re.M)
    text = re.sub(r"^(\s*extra-lib-dirs:\s*)(\S+)\s*$", lambda m: f"{m.group(1)}{m.group(2)}, {lib_dir}", text, count=1, flags=re.M)
    return text


def patch_liquid(text):
    text = insert_after(text, r'^foreign import ccall unsafe "agc_crcf_destroy" c_agc_crcf_destroy\n\s+:: Agc -> IO \(\)\n', AGC_IMPORT, "agc_crcf_destroy import")
    text = replace_between(text, r"^agcExecuteBlock agc px n py = do$", r"^agcCreate :: ", AGC_BODY, "agcExecuteBlock")
    text = insert_before(text, r"^firpfbchCreate :: ", PFB_IMPORT, "firpfbchCreate")
    text = replace_between(text, r"^firpfbchChan \(fb, nco, nch\) a = do$", r"^firpfbchChannelizer ::", PFB_BODY, "firpfbchChan")
    return text


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("checkout")
    ap.add_argument("--lib-dir", default="/opt/composable-sdr_b200")
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args()
    jobs = [("composable-sdr.cabal", lambda t: patch_cabal(t, args.lib_dir)),
            (os.path.join("src", "ComposableSDR", "Liquid.chs"), patch_liquid)]
    done = []
    for rel, fn in jobs:                       # every anchor is checked before anything is written
        path = os.path.join(args.checkout, rel)
        old = open(path).read()
        done.append((rel, path, old, fn(old)))
    for rel, path, old, new in done:
        sys.stdout.writelines(difflib.unified_diff(old.splitlines(True), new.splitlines(True), "a/" + rel, "b/" + rel, n=1))
        if not args.dry_run:
            open(path, "w").write(new)


if __name__ == "__main__":
    main()
