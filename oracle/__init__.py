"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's progressive path (the parity checker).

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
"""
