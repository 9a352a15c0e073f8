"""``watch.Watch().stream(list_fn, ...)``: ADDED / MODIFIED / DELETED events by polling the list function (the daemon's REST
API has no chunked watch); ends after ``timeout_seconds`` or ``stop()``."""
import time


class Watch:
    def __init__(self):
        self._stop = False

    def stop(self) -> None:
        self._stop = True

    def stream(self, func, *args, timeout_seconds=None, _poll=0.2, **kwargs):
        seen = {}
        deadline = time.time() + timeout_seconds if timeout_seconds else None
        while not self._stop and (deadline is None or time.time() < deadline):
            res = func(*args, **kwargs)
            items = res.get("items", []) if isinstance(res, dict) else getattr(res, "items", [])
            cur = {}
            for o in items:
                md = o["metadata"]
                key = (md.get("namespace", ""), md["name"])
                cur[key] = o
                rv = md.get("resourceVersion")
                if key not in seen:
                    seen[key] = rv
                    yield {"type": "ADDED", "object": o, "raw_object": o}
                elif seen[key] != rv:
                    seen[key] = rv
                    yield {"type": "MODIFIED", "object": o, "raw_object": o}
                if self._stop:
                    return
            for key in [k for k in seen if k not in cur]:
                del seen[key]
                yield {"type": "DELETED", "object": {"metadata": {"namespace": key[0], "name": key[1]}}, "raw_object": None}
            time.sleep(_poll)
