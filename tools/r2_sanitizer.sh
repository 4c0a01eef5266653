#!/bin/bash
# compute-sanitizer sweep over the small-picture GPU tests (memcheck, racecheck, initcheck, synccheck); output -> profiles/r02_sanitizer.txt
T="tests/test_gpu_synthetic.py tests/test_gpu_bands.py tests/test_gpu_sao_stats.py tests/test_gpu_alf_stats.py tests/test_gpu_postfilter.py tests/test_gpu_parity.py"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -m gpu -x -q -k "not 3840 and not 4k" 2>&1 | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_sao_stats.py tests/test_gpu_alf_stats.py tests/test_gpu_postfilter.py tests/test_gpu_parity.py -m gpu -x -q -k "not 1920 and not 4k and not 3840" 2>&1 | tail -3
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_postfilter.py -m gpu -x -q -k "not 1920 and not 3840" 2>&1 | tail -3
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sao_stats.py tests/test_gpu_alf_stats.py -m gpu -x -q -k "not 1920 and not 4k" 2>&1 | tail -3
