python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --particles 1e6 --steps 3 --warmup 3 --cpu-nx 30 2>&1 | tail -3
python bench.py --particles 16e6 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -3
