"""The non-residual model families tf_cnn_benchmarks accepts for ``--model=`` (the reference's headline job passes
``--model=resnet101``, examples/v2beta1/tensorflow-benchmarks/tensorflow-benchmarks.yaml:38-42; its README tells users
to edit that flag): ``trivial``, ``lenet``, ``alexnet``, ``vgg11/16/19``. Plain PyTorch modules; convolutions and GEMMs
run in cuDNN / cuBLAS, gradients and the optimizer go through the fused window path like every other model."""
import torch
import torch.nn as nn


class Trivial(nn.Module):
    """One hidden layer on the flattened image: the communication-bound smoke model (about 150 M parameters per
    1000 classes at 224x224 would be too much; like tf_cnn's trivial model the hidden layer is one unit wide and the
    classifier 4096 wide)."""

    def __init__(self, num_classes: int = 1000, image_size: int = 224):
        super().__init__()
        self.fc1 = nn.Linear(3 * image_size * image_size, 1)
        self.fc2 = nn.Linear(1, 4096)
        self.fc3 = nn.Linear(4096, num_classes)

    def forward(self, x):
        x = torch.flatten(x, 1)
        return self.fc3(torch.relu(self.fc2(torch.relu(self.fc1(x)))))


class LeNet(nn.Module):
    def __init__(self, num_classes: int = 1000, in_channels: int = 3):
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv2d(in_channels, 32, 5, padding=2), nn.ReLU(inplace=True), nn.MaxPool2d(2),
            nn.Conv2d(32, 64, 5, padding=2), nn.ReLU(inplace=True), nn.MaxPool2d(2),
            nn.AdaptiveAvgPool2d(7))
        self.classifier = nn.Sequential(nn.Linear(64 * 7 * 7, 512), nn.ReLU(inplace=True), nn.Linear(512, num_classes))

    def forward(self, x):
        return self.classifier(torch.flatten(self.features(x), 1))


class AlexNet(nn.Module):
    def __init__(self, num_classes: int = 1000):
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv2d(3, 64, 11, stride=4, padding=2), nn.ReLU(inplace=True), nn.MaxPool2d(3, 2),
            nn.Conv2d(64, 192, 5, padding=2), nn.ReLU(inplace=True), nn.MaxPool2d(3, 2),
            nn.Conv2d(192, 384, 3, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(384, 256, 3, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(256, 256, 3, padding=1), nn.ReLU(inplace=True), nn.MaxPool2d(3, 2),
            nn.AdaptiveAvgPool2d(6))
        self.classifier = nn.Sequential(
            nn.Linear(256 * 36, 4096), nn.ReLU(inplace=True),
            nn.Linear(4096, 4096), nn.ReLU(inplace=True),
            nn.Linear(4096, num_classes))

    def forward(self, x):
        return self.classifier(torch.flatten(self.features(x), 1))


_VGG = {
    11: (1, 1, 2, 2, 2),
    16: (2, 2, 3, 3, 3),
    19: (2, 2, 4, 4, 4),
}


class VGG(nn.Module):
    def __init__(self, depth: int = 16, num_classes: int = 1000):
        super().__init__()
        layers, cin = [], 3
        for reps, cout in zip(_VGG[depth], (64, 128, 256, 512, 512)):
            for _ in range(reps):
                layers += [nn.Conv2d(cin, cout, 3, padding=1), nn.ReLU(inplace=True)]
                cin = cout
            layers.append(nn.MaxPool2d(2))
        layers.append(nn.AdaptiveAvgPool2d(7))
        self.features = nn.Sequential(*layers)
        self.classifier = nn.Sequential(
            nn.Linear(512 * 49, 4096), nn.ReLU(inplace=True),
            nn.Linear(4096, 4096), nn.ReLU(inplace=True),
            nn.Linear(4096, num_classes))

    def forward(self, x):
        return self.classifier(torch.flatten(self.features(x), 1))


def trivial(**kw):
    return Trivial(**kw)


def lenet(**kw):
    return LeNet(**kw)


def alexnet(**kw):
    return AlexNet(**kw)


def vgg11(**kw):
    return VGG(11, **kw)


def vgg16(**kw):
    return VGG(16, **kw)


def vgg19(**kw):
    return VGG(19, **kw)
