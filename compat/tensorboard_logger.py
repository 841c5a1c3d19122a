"""Shim of tensorboard_logger (reference train.py:44-45): scalars go to torch's SummaryWriter when
available, otherwise they are dropped."""
_writer = None


def configure(path, flush_secs=2):
    global _writer
    try:
        from torch.utils.tensorboard import SummaryWriter
        _writer = SummaryWriter(path, flush_secs=flush_secs)
    except Exception:
        _writer = None


def log_value(name, value, step):
    if _writer is not None:
        _writer.add_scalar(name, value, step)
