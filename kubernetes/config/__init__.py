"""``config.load_kube_config()`` / ``load_incluster_config()``: choose the daemon instead of a kubeconfig.
Order: explicit ``host=`` argument, ``$MPIJOB_SERVER``, ``~/.mpijob/server`` (one line ``host:port``), 127.0.0.1:8087."""
import os

_DEFAULT = "127.0.0.1:8087"
_state = {"host": None}


class ConfigException(Exception):
    pass


def _resolve(host=None) -> str:
    if host:
        return host
    if os.environ.get("MPIJOB_SERVER"):
        return os.environ["MPIJOB_SERVER"]
    path = os.path.expanduser("~/.mpijob/server")
    if os.path.exists(path):
        with open(path) as f:
            line = f.read().strip()
        if line:
            return line
    return _DEFAULT


def load_kube_config(config_file=None, context=None, client_configuration=None, persist_config=True, host=None) -> None:
    h = _resolve(host)
    _state["host"] = h if h.startswith("http") else "http://" + h
    if client_configuration is not None:
        client_configuration.host = _state["host"]


def load_incluster_config(client_configuration=None) -> None:
    load_kube_config(client_configuration=client_configuration)


def load_config(**kwargs) -> None:
    load_kube_config(**kwargs)


def current_host() -> str:
    if _state["host"] is None:
        load_kube_config()
    return _state["host"]
