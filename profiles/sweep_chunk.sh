for c in 2 4 8 16; do echo chunk=$c; FO_FWD_CHUNK=$c python profiles/fwd_variant_probe.py; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
