"""Drop-in for the reference ``operators`` package (operators/segment_reduction): the
``_ext.segment_reduction`` native module with the four names the reference python side calls,
the autograd ``Function`` and the ``nn.Module`` wrapper."""
