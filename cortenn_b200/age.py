"""Graph builders: one function per entry of the reference's op table
(llo/cfg/llo.json:125-531), second overloads suffixed 0 as pybinder does
(pybinder/pyapi_plugin.py:41-54), plus the `Tensor` class."""
from ._C import age as _age

for _name in dir(_age):
    if not _name.startswith("_"):
        globals()[_name] = getattr(_age, _name)
del _name
