"""Single-rank stand-in for mpi4py, used ONLY by tools/make_case.py to run the
reference assembler (bin/assemble.py) unmodified in a container without MPI.

It implements exactly the calls made at /root/reference/bin/assemble.py:46-48,
66-67, 548, 574-576, 596, 1128, 1156-1158, 1176 for a communicator of size 1.
Test infrastructure; never imported by the product path.
"""
from . import MPI  # noqa: F401
