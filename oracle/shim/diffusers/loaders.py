class UNet2DConditionLoadersMixin:
    pass


class LoraLoaderMixin:
    pass


class TextualInversionLoaderMixin:
    pass
