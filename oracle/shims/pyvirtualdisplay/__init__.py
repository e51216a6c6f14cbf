"""pyvirtualdisplay stand-in (oracle/test infrastructure only).

rltoolkit/tensorboard_logger.py:11,371,395 uses `Display` as a context manager around video
recording; nothing on the hot path needs it.
"""


class Display:
    def __init__(self, *args, **kwargs):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def start(self):
        return self

    def stop(self):
        return self
