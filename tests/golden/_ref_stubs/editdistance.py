"""Stand-in for the `editdistance` package (imported by the reference's eval_script.py, unused on the IoU path)."""


def eval(a, b):          # noqa: A001 - the package's own name for it
    raise NotImplementedError("not needed for the IoU golden vectors")
