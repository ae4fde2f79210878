class IPAdapterMaskProcessor:
    pass


PipelineImageInput = object  # a typing alias upstream
