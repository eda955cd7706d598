"""network registry (deepxube/factories/nnet_factory.py)."""
from deepxube_b200.base.factory import Factory

deepxube_nnet_factory: Factory = Factory("DeepXubeNNet")
