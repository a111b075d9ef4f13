timeout 600 python -m pytest tests -m gpu -x -q -k "shard_hint or models" 2>&1 | tail -4
