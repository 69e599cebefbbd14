"""Command-line front end (`portcullis <mode> ...`): mode dispatch, help, option errors, and the no-GPU failure mode.  Mirrors the
behaviour of src/portcullis.cc:111-127, 406-517 for the modes this build provides.  CPU only."""
import os
import subprocess

import pytest

from conftest import GOLDEN, make_prep

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "portcullis_b200", "bin", "portcullis")


def run(*args):
    return subprocess.run([EXE, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def test_mode_dispatch_and_help():
    p = run()
    assert p.returncode == 1 and "portcullis junc" in p.stderr and "portcullis prep" in p.stderr and "portcullis bamfilt" in p.stderr
    for mode, marker in (("junc", "Usage: portcullis junc"), ("JUNC", "Usage: portcullis junc"), ("analyse", "Usage: portcullis junc"),
                         ("prep", "Usage: portcullis prep"), ("Prepare", "Usage: portcullis prep"), ("bamfilt", "Usage: portcullis bamfilt")):
        p = run(mode, "--help")
        assert p.returncode == 1 and marker in p.stdout, mode
    p = run("filt")
    assert p.returncode == 1 and "not provided by this build" in p.stderr
    assert run("--version").stdout.startswith("portcullis 1.2.4")


def test_option_errors():
    p = run("junc", "--no_such_option", "x")
    assert p.returncode == 1 and "unrecognised option" in p.stderr
    p = run("junc", "-t")
    assert p.returncode == 1 and "required argument" in p.stderr
    p = run("junc", "a", "b")
    assert p.returncode == 1 and "too many positional" in p.stderr
    p = run("junc", "--orientation", "sideways", "/nonexistent")
    assert p.returncode != 0
    p = run("junc", "/nonexistent/prep")
    assert p.returncode == 4 and "Could not find prepared BAM file" in p.stderr
    p = run("bamfilt", "--clip_mode", "medium", "a.tab", "b.bam")
    assert p.returncode == 1 and "clip mode" in p.stderr
    p = run("prep", "-o", "/tmp/pj_cli_none", "/nonexistent.fa", "x.bam")
    assert p.returncode == 4 and "Could not find genome file" in p.stderr


def test_junc_without_a_gpu_fails_loudly(tmp_path):
    """There is no CPU fallback: on a machine without a CUDA device `junc` stops with the library's message, exit code 4."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    p = run("junc", "-o", str(tmp_path / "o" / "p"), make_prep(tmp_path, "kat"))
    assert p.returncode == 4 and "no CPU fallback" in p.stderr
    assert not os.path.exists(str(tmp_path / "o" / "p.junctions.tab"))
