"""``horovod.common``: exceptions and small helpers scripts import from here."""
