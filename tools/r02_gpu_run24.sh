timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "full_geometry or no_freeze" -s 2>&1 | grep -E "PARITY|passed|failed|Error|assert" | head -30
