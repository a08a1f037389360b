"""TEST INFRASTRUCTURE ONLY -- `Data`/`Dataset` stand-ins (attribute bag + `__inc__` + `.to`)."""
import torch


class Data:
    def __init__(self, *args, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def __inc__(self, key, value, *args, **kwargs):
        return 0

    def to(self, device, *args, **kwargs):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


class Dataset:
    def __init__(self, *args, **kwargs):
        pass
