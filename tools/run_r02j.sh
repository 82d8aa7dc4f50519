timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "config2 or stripe or split or ragged or scan or match" 2>&1 | tail -2
SHAPES=3 ROWS=512,384,256,192,128,96,64,48,32,16 timeout 300 python tools/sweep_match3.py
