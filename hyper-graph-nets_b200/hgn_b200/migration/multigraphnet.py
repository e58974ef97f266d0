"""Mirror of src/migration/multigraphnet.py: the base block on the one-hot tagged merged graph."""
from .graphnet import GraphNet


class MultiGraphNet(GraphNet):
    """Multi-Edge and Multi-Node Interaction Network with residual connections."""
