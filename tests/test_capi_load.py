"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/rsb200.h
declares, answers the calls that need no GPU, and refuses (never falls back) without a device."""
import ctypes as C
import os
import subprocess

import pytest


def test_library_exports_every_declared_symbol(rsb):
    lib = rsb.load_library()
    declared = rsb.declared_symbols()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    # and nothing the header does not declare leaks out with the rs_ prefix
    out = subprocess.run(["nm", "-D", "--defined-only", rsb.library_path()], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T rs_" in l)
    assert exported == declared


def test_status_strings_follow_operation_status(rsb):
    # ordinals of the reference's enum class OperationStatus (src/OperationStatus.h:4-17)
    names = ["OK", "HOST_BUFFERS_FAILED", "INITIALIZATION_FAILED", "DATA_UPLOAD_FAILED", "CALCULATION_FAILED",
             "DATA_DOWNLOAD_FAILED", "CLEANUP_FAILED", "RESIZE_FAILED", "KERNEL_CREATION_FAILED",
             "PROGRAM_CREATION_FAILED", "NO_SOURCE_FOUND", "LOADING_SOURCE_FAILED"]
    assert [rsb.status_string(i) for i in range(12)] == names
    assert rsb.status_string(99) == "UNKNOWN"


def test_round_up_keys_is_resize(rsb):
    # RadixSortGPU::Resize: next multiple of _NUM_GROUPS*_NUM_ITEMS_PER_GROUP = 1024
    assert [rsb.round_up_keys(n) for n in (0, 1, 1023, 1024, 1025, 1 << 20, (1 << 20) + 1)] == \
        [0, 1024, 1024, 1024, 2048, 1 << 20, (1 << 20) + 1024]


def test_no_device_means_loud_failure(rsb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rsb.RsError) as e:
        rsb.Context(0)
    assert e.value.status == 2  # INITIALIZATION_FAILED; there is no CPU fallback behind the ABI


def test_header_is_plain_c(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "t.c"
    src.write_text('#include "rsb200.h"\nint main(void){ rs_runtimes r; (void)r; return RS_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-c", str(src), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_product_never_touches_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "radix-sort_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("oracle (", ""), os.path.join(dirpath, f)


def test_bench_reference_arm_line():
    # bench.py --impl reference: the reference's own CPU sort on a bounded sample, contract keys present
    import json
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Gkeys/s" and line["higher_is_better"] is True
    assert line["metric"] == "Gkeys/s sorting 2^28 uint32 keys" and 0 < line["value"] < 1.0
    assert line["cpu_baseline"]["cores"] == 1 and line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["e2e"] == {"value": line["value"], "unit": "Gkeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_merge_tile_order_covers_every_tile_once():
    # the region-strided order in which the merge CTAs can walk a batch (csrc/rs_merge_order.h, the REGIONS
    # instantiation of the tile kernel; opt-in until it has run on hardware), checked on the host:
    # the C++ function itself (bijection, few empty slots) and the kernel's seek/advance/locate walk for many job shapes
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "tests", "cpp", "merge_order_check")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(root, "tests", "cpp"), "merge_order_check"], check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "merge order ok" in out.stdout, out.stdout + out.stderr
    import sys
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "merge_order_emulation.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "every tile exactly once" in out.stdout, out.stdout + out.stderr
