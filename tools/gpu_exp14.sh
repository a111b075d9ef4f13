timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "site-packages" | tail -25
