from .channels_strategies import one_channel_collate_fn, OneChannelCollator  # noqa: F401
