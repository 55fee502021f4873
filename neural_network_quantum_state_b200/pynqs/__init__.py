"""Drop-in for the reference's `pynqs` package (python/pynqs/__init__.py): `from pynqs import sampler`.

Use it either as `neural_network_quantum_state_b200.pynqs` or, to keep reference scripts (python/meas_*.py) unchanged, put
this package's parent directory ... /neural_network_quantum_state_b200 on PYTHONPATH so that `import pynqs` resolves here.
"""
__all__ = ["sampler"]
