"""Exception types of the Horovod-compatible front-end (same names as horovod.common.exceptions)."""


class HorovodInternalError(RuntimeError):
    """A collective failed (mismatched submissions, a dead peer, shutdown). Elastic training catches it and restores the
    last committed state."""

    def __init__(self, msg: str = "", code: int = 0):
        super().__init__(msg)
        self.code = code


class HostsUpdatedInterrupt(RuntimeError):
    """Raised inside a training function when the host set changed."""

    def __init__(self, msg: str = "", skip_sync: bool = False):
        super().__init__(msg)
        self.skip_sync = skip_sync
