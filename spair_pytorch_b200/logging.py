"""Logging hooks — empty in the reference too (spair/logging.py:2-9); kept for import compatibility."""


def log():
    pass


def record_scalar(t, name, group):
    pass


def record_image():
    pass
