"""Harness code behind bench.py and __graft_entry__.smoke(): workloads, CPU arms, smoke step.
Not part of the product package (mem_b200/); the only code besides tests/ that may import oracle/."""
