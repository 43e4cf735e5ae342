"""TEST INFRASTRUCTURE: CPU oracle of the hypo_tremor_mcmc hot path (see htm_oracle.hpp).
Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs."""
