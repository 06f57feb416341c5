python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --particles 8e6 --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1
