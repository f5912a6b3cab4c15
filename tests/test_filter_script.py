"""CPU: vadc_b200/filter_script (timestamps -> ffmpeg aselect expression) against the reference's unmodified
filter_script.c built by oracle/Makefile (oracle/_ref/filter_script_ref), byte for byte; plus the multi-file extension."""
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import ROOT

TOOL = os.path.join(ROOT, "vadc_b200", "filter_script")
REF = os.path.join(ROOT, "oracle", "_ref", "filter_script_ref")


def run(exe, text):
    r = subprocess.run([exe], input=text.encode(), capture_output=True, timeout=30)
    assert r.returncode == 0
    return r.stdout.decode()


def _cases():
    rng = np.random.default_rng(3)
    t = np.cumsum(rng.uniform(0.05, 7.0, 400)).astype(np.float32)
    seconds = "".join("%.2f,%.2f\n" % (a, b) for a, b in zip(t[0::2], t[1::2]))            # vadc.c:244-250
    centi = "".join("%d,%d\n" % (int(a * 100), int(b * 100)) for a, b in zip(t[0::2], t[1::2]))   # vadc.c:251-256
    return {"seconds": seconds, "centiseconds": centi, "one": "0.10,0.58\n", "empty": "", "no_final_newline": "1.00,2.00\n3.5,4.25",
            "ten_hours": "35990.02,35999.90\n"}


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/filter_script_ref not built (no /root/reference here)")
@pytest.mark.parametrize("name", sorted(_cases()))
def test_same_bytes_as_reference_tool(name):
    text = _cases()[name]
    assert run(TOOL, text) == run(REF, text)


def test_known_answer():
    assert run(TOOL, "0.50,1.25\n3.00,4.10\n") == \
        "asetpts=N/SR/TB, aselect='between(t,0.500000,1.250000)+between(t,3.000000,4.100000)', asetpts=N/SR/TB"
    assert run(TOOL, "") == "asetpts=N/SR/TB, aselect='', asetpts=N/SR/TB"
    # garbage ends the list instead of spinning forever (the reference's scanf loop never terminates on it)
    assert run(TOOL, "1.0,2.0\nhello\n3.0,4.0\n") == "asetpts=N/SR/TB, aselect='between(t,1.000000,2.000000)', asetpts=N/SR/TB"


def test_multi_file_listing_of_the_cli():
    text = "# a.s16le\n0.10,0.58\n1.00,2.00\n# b.s16le\n# c.s16le\n5.00,6.50\n"
    want = ("# a.s16le\nasetpts=N/SR/TB, aselect='between(t,0.100000,0.580000)+between(t,1.000000,2.000000)', asetpts=N/SR/TB\n"
            "# b.s16le\nasetpts=N/SR/TB, aselect='', asetpts=N/SR/TB\n"
            "# c.s16le\nasetpts=N/SR/TB, aselect='between(t,5.000000,6.500000)', asetpts=N/SR/TB\n")
    assert run(TOOL, text) == want
