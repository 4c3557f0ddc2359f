"""Route `print` through a tqdm bar while a run is active (reference Algorithms/utils/tqdm_write_stream.py)."""
import sys


class _TqdmWriter:
    def __init__(self, pbar):
        self._pbar = pbar
        self._buf = ""

    def write(self, text):
        self._buf += text
        while "\n" in self._buf:
            line, self._buf = self._buf.split("\n", 1)
            self._pbar.write(line)

    def flush(self):
        if self._buf:
            self._pbar.write(self._buf)
            self._buf = ""


def redirect_stdout_to_tqdm(pbar):
    original = sys.stdout
    sys.stdout = _TqdmWriter(pbar)
    return original


def restore_stdout(original):
    if original is not None:
        try:
            sys.stdout.flush()
        except Exception:
            pass
        sys.stdout = original
