"""CPU oracle of the tracer hot path -- TEST INFRASTRUCTURE (see polaris_oracle.cpp header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs.  Nothing under polaris_b200/ imports this package.
"""
