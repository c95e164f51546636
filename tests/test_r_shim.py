"""The R package side (r/): the `.Call` shim must compile against the R C API (a stub of it -- R is not installed here) and
against include/ruvnb.h, i.e. every entry point it binds exists with the argument list it uses; the R wrappers keep the
reference's exported signatures."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_compiles_against_the_header():
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    r = subprocess.run([cc, "-std=c99", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-Wno-cast-function-type",   # (DL_FUNC casts: every R package has them)
                       
                        "-I", os.path.join(ROOT, "tests", "stubs"), "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r", "src", "ruvnb_shim.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_binds_only_declared_symbols_and_registers_its_routines():
    shim = open(os.path.join(ROOT, "r", "src", "ruvnb_shim.c")).read()
    hdr = open(os.path.join(ROOT, "include", "ruvnb.h")).read()
    used = set(re.findall(r"\b(ruvnb_[a-z_A-Z0-9]+)\s*\(", shim)) - set(re.findall(r"\b(ruvnb_R_[a-z_A-Z0-9]+)", shim))
    declared = set(re.findall(r"\b(ruvnb_[a-z_A-Z0-9]+)\s*\(", hdr))
    assert used and used <= declared, used - declared
    defined = set(re.findall(r"^SEXP (ruvnb_R_[a-z_A-Z0-9]+)\(", shim, flags=re.M))
    registered = set(re.findall(r'\{"(ruvnb_R_[a-z_A-Z0-9]+)"', shim))
    assert defined == registered
    wrappers = open(os.path.join(ROOT, "r", "R", "ruvIIInb_gpu.R")).read()
    called = set(re.findall(r'\.Call\("(ruvnb_R_[a-z_A-Z0-9]+)"', wrappers))
    assert called and called <= registered


def test_r_wrappers_keep_the_reference_signatures():
    """Argument names of the reference's exported functions (man/*.Rd) must all be accepted."""
    w = open(os.path.join(ROOT, "r", "R", "ruvIIInb_gpu.R")).read()
    want = {
        "ruvIII.nb": "Y M ctl k robust ortho.W lambda.a lambda.b batch step.fac inner.maxit outer.maxit ncores use.pseudosample nc.pool batch.disp zeroinf",
        "fastruvIII.nb": "Y M cData ctl k robust ortho.W lambda.a lambda.b batch step.fac inner.maxit outer.maxit ncores use.pseudosample nc.pool batch.disp pCells.touse block.size",
        "estimateDisp.par": "y design group lib.size offset prior.df trend.method tagwise span min.row.sum grid.length grid.range robust winsor.tail.p tol weights BPPARAM",
        "get.res": "out type batch block.size",
        "makeSCE": "obj cData batch assays",
    }
    for fn, args in want.items():
        m = re.search(re.escape(fn) + r" <- function\((.*?)\)\s*\{", w, flags=re.S)
        assert m, fn
        have = set(a.split("=")[0].strip() for a in re.split(r",(?![^()]*\))", m.group(1)))
        assert set(args.split()) <= have, (fn, set(args.split()) - have)
    for text in ("Less than 1% of cells have known annotation", "of the M matrix has zero sum", "Batch variable is not a numeric vector",
                 "Incorrect length of group.", "Incorrect length of lib.size.", "'cData' must be a user-specified data frame.",
                 "Wrong names of corrected assays data was specified"):
        assert text in w
