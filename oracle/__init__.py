"""TEST INFRASTRUCTURE ONLY.

CPU fp32 restatements of the reference's per-block algorithms (oracle.restate) and the loader
that imports the real reference modules in the authoring container (oracle.ref_loader).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (jittor-mlp_b200/) never does and has no CPU path.
"""
