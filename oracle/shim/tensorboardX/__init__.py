"""TEST INFRASTRUCTURE ONLY (oracle shim): no-op SummaryWriter (reference utils/saver.py:6,32)."""


class SummaryWriter(object):
    def __init__(self, *a, **k):
        pass

    def add_text(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def close(self):
        pass
