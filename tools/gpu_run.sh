python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python tools/show.py /dev/stdin | head -1
