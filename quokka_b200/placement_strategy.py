"""Placement strategies -- same names / arguments as pyquokka/placement_strategy.py:1-36.  A "node" of
the reference is a GPU (rank) here; `channels_per_node` > 1 is accepted and treated as 1 (one channel
per GPU: the kernels already use the whole device)."""


class PlacementStrategy:
    def __init__(self) -> None:
        pass


class SingleChannelStrategy(PlacementStrategy):
    """One channel in total (rank 0): final aggregates, top-k."""


class CustomChannelsStrategy(PlacementStrategy):
    def __init__(self, channels) -> None:
        super().__init__()
        self.channels_per_node = channels


class DatasetStrategy(PlacementStrategy):
    def __init__(self, total_channels) -> None:
        super().__init__()
        self.total_channels = total_channels


class TaggedCustomChannelsStrategy(PlacementStrategy):
    def __init__(self, channels, tag) -> None:
        super().__init__()
        self.channels_per_node = channels
        self.tag = tag
