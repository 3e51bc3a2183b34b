for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed|^E  " | head -5; done
