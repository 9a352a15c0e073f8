"""The MNIST conv net of the reference's Horovod example
(examples/v2beta1/horovod/tensorflow_mnist.py:38-73): two 5x5 conv layers
(32, 64 feature maps) with 2x2 max-pooling, FC-1024 with dropout, FC-10."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class MnistConvNet(nn.Module):
    def __init__(self, dropout: float = 0.5):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 32, 5, padding=2)
        self.conv2 = nn.Conv2d(32, 64, 5, padding=2)
        self.fc1 = nn.Linear(7 * 7 * 64, 1024)
        self.fc2 = nn.Linear(1024, 10)
        self.drop = nn.Dropout(dropout)

    def forward(self, x):
        x = x.view(-1, 1, 28, 28)
        x = F.max_pool2d(F.relu(self.conv1(x)), 2)
        x = F.max_pool2d(F.relu(self.conv2(x)), 2)
        x = self.drop(F.relu(self.fc1(torch.flatten(x, 1))))
        return self.fc2(x)
