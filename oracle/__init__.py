"""CPU oracle for the vqgan-training hot path.

TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is imported by the product package
(vqgan-training_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may use it, and there only as the checker / CPU baseline.

The oracle is a plain-PyTorch fp32 functional restatement of the reference's modules
(/root/reference ae.py, utils.py, vae_trainer.py; every function cites the lines it follows),
operating directly on reference-format state_dicts. It is PINNED against the reference itself:
oracle/make_golden.py imports the unmodified reference in the build container, runs it on seeded
weights/inputs (oracle/seeded.py) and commits the outputs under tests/golden/; tests/test_oracle_golden.py
checks the restatement against those vectors on every CPU test run. The reference ships no golden
vectors or tests of its own, so these fixtures are the pin.
"""
