"""Client machinery: object store, typed clientset, fake, informers, listers,
apply configurations."""
from .clientset import Action, Clientset, FakeClientset, KubeClient, MPIJobInterface, ResourceClient  # noqa: F401
from .errors import ApiError, is_already_exists, is_conflict, is_not_found  # noqa: F401
from .informers import Indexer, Lister, MPIJobLister, SharedIndexInformer, SharedInformerFactory  # noqa: F401
from .store import ADDED, DELETED, MODIFIED, ObjectStore  # noqa: F401
