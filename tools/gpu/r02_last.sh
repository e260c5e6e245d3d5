#!/bin/bash
timeout 600 python -m pytest tests/test_scattering2d_gpu.py tests/test_autograd2d_gpu.py tests/test_filters_gpu.py -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 200 python tools/kbench.py c2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 %.3f ms %.0f img/s chk %.10e'%(d['ms_median'], d['img_per_s'], d['checksum']))"
