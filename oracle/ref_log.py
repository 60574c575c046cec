"""The reference's per-epoch log / print contract (main.py:190-201, 218-244) executed FROM ITS OWN SOURCE TEXT.
TEST INFRASTRUCTURE.  main.py cannot be imported (it needs visdom and runs a whole experiment at import), so the two
blocks are cut out of the unmodified file by line content, dedented and exec'd in a sandbox namespace:
``epoch_update(log, stat, epoch_time) -> printed lines`` runs exactly the statements the reference runs at the end
of an epoch.  Used by tests/test_log_contract.py (differential) and oracle/gen_golden.py (fixture)."""
import io
import os
import textwrap
from contextlib import redirect_stdout

import numpy as np

from . import ref_shims


def _source():
    with open(os.path.join(ref_shims.REF_ROOT, "main.py")) as f:
        return f.read().split("\n")


def _block(lines, first_startswith, last_startswith):
    i0 = next(i for i, ln in enumerate(lines) if ln.strip().startswith(first_startswith))
    i1 = max(i for i, ln in enumerate(lines) if ln.strip().startswith(last_startswith))
    return textwrap.dedent("\n".join(lines[i0:i1 + 1]))


def make_log():
    """log = dict(); log['epoch'] = LogField(...) ... (main.py:190-201) with the reference's utils.LogField."""
    ref_shims.install()
    import utils as ref_utils
    ns = dict(LogField=ref_utils.LogField)
    exec(_block(_source(), "log = dict()", "log['entropy']"), ns)
    return ns["log"]


def epoch_update(log, stat, epoch_time):
    """main.py:218-244 on (log, stat): normalises stat in place, appends to the log, returns the printed lines."""
    src = _block(_source(), "epoch = len(log['epoch'].data) + 1", "print('Enemy-Comm")
    ns = dict(log=log, stat=stat, epoch_time=epoch_time, np=np)
    buf = io.StringIO()
    with redirect_stdout(buf):
        exec(src, ns)
    return buf.getvalue().rstrip("\n").split("\n")
