python tools/mesh_ambiguity.py 0.05 2>&1 | tail -2
python tools/mesh_ambiguity.py 0.1 2>&1 | tail -1
