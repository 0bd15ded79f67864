"""profiles/front_kernel_static.json (what bench.py reports as roofline.traffic / executed_ops_per_unit) must describe the
library that is built: its recorded hot-loop instruction count is recomputed from the current .so's SASS."""
import json
import os
import shutil
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs cuobjdump")
def test_static_json_matches_the_built_library():
    import front_static
    if not os.path.exists(front_static.LIB):
        pytest.skip("library not built")
    rec = json.load(open(front_static.OUT))
    now = front_static.sass_counts()
    for kernel in ("ddc_front_tc_kernel", "ddc_front_bt_kernel", "ddc_front_kernel"):
        assert now[kernel]["hot_loop_sass_instructions"] == rec[kernel]["hot_loop_sass_instructions"], \
            "%s changed: re-profile and run tools/front_static.py update" % kernel
        # the ncu-measured executed instructions per unit must sit just above the hot loop's static count
        static = now[kernel]["hot_loop_sass_instructions_per_unit"]
        assert static <= rec[kernel]["sass_instructions_per_unit"] <= 1.12 * static
        assert rec[kernel]["traffic_bytes_per_launch"] > 1e8 and "source" in rec[kernel]
    assert 0.11 < rec["ddc_front_tc_kernel"]["lsu_wavefronts_per_unit"] < 0.2      # >= E[max bank load] / 32 of the table gather
