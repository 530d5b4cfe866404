"""`llo.llo` surface of the reference (llo/python/llo.cpp:156-215): variable,
Variable.assign, evaluate, derive, seed, print_graph — evaluated on the B200."""
from ._C import llo as _llo

for _name in dir(_llo):
    if not _name.startswith("_"):
        globals()[_name] = getattr(_llo, _name)
del _name
