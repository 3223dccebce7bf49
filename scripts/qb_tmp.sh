timeout 900 python -m pytest tests -q -x -m gpu --timeout 300 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
