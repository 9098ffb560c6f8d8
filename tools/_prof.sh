for k in 3; do timeout 60 python tools/run_glyph.py gauss_s16 5000000 $k | tail -1; timeout 60 python tools/run_glyph.py gauss_s4 5000000 $k | tail -1; done
timeout 200 python -m pytest tests/test_parity_gpu.py -x -q -k "gaussian_vs_oracle or gather_edge" 2>&1 | tail -3
