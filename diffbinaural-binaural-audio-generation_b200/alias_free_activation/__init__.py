"""`alias_free_activation` as the reference imports it (reference tree: BigVGAN/bigvgan.py:19, :96).

Put this directory's parent AHEAD of the reference's BigVGAN/ on sys.path:
  * `alias_free_activation.cuda.activation1d.Activation1d` -> the fused B200 kernel;
  * `alias_free_activation.torch.{act,filter,resample}` -> the reference's own flat files
    (BigVGAN/alias_free_activation/*.py), loaded unmodified under the names they expect -- the
    reference tree forgot that sub-package.  Nothing of it is re-implemented here.
"""
