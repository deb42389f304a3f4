timeout 600 python scripts/batch_streams.py 22 24 2>&1 | tail -5
timeout 600 python scripts/batch_streams.py 20 32 2>&1 | tail -5
