"""``watch.Watch().stream(list_fn, ...)``: ADDED / MODIFIED / DELETED events. For the list functions of this package's
``CustomObjectsApi`` / ``CoreV1Api`` the daemon's chunked ``?watch=true`` stream is used; any other callable is polled. Ends
after ``timeout_seconds`` or ``stop()``."""
import time


class Watch:
    def __init__(self):
        self._stop = False

    def stop(self) -> None:
        self._stop = True

    def stream(self, func, *args, timeout_seconds=None, _poll=0.2, **kwargs):
        owner = getattr(func, "__self__", None)
        c = getattr(owner, "_c", None)
        streams = {"list_namespaced_custom_object": ("mpijobs", 2), "list_cluster_custom_object": ("mpijobs", None),
                   "list_namespaced_pod": ("pods", 0), "list_namespaced_event": ("events", 0)}
        if c is not None and getattr(func, "__name__", "") in streams:
            resource, ns_arg = streams[func.__name__]
            ns = kwargs.get("namespace", args[ns_arg] if ns_arg is not None and len(args) > ns_arg else None)
            for ev in c.watch(resource, ns, timeout=timeout_seconds or 300.0, label_selector=kwargs.get("label_selector")):
                if self._stop:
                    return
                yield {"type": ev["type"], "object": ev["object"], "raw_object": ev["object"]}
            return
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
