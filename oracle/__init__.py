"""TEST INFRASTRUCTURE: CPU oracle (numpy/scipy) for the GP scoring-rule hot path.

Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs.  See gp_oracle.py for the header.
"""
