"""Import shim with the reference's module name: ``torch.save(model)`` in the reference pickles its classes as
``cVAE.<Class>``, so a ``cVAE_model.pkl`` written by the reference resolves to the drop-in classes
here when this directory is on ``sys.path`` (the multimodal_kfold_* programs run from it)."""
from multi_modal_normative_modeling_b200.cVAE import (  # noqa: F401
    DEVICE, Decoder, Discriminator, Encoder, cVAE, cVAE_multimodal, cVAE_multimodal_regression, compute_ll, fuse_latent, mmJSD, mvtCAE)
# the reference's cVAE module exports the END-TO-END SUPERVISED model (v2, cVAE.py:2021-2207) under this name; the different
# class of the same name defined inside multimodal_kfold_cvae_nmmlp.py is multi_modal_normative_modeling_b200.cVAE's
from multi_modal_normative_modeling_b200.e2e import Classifier, cVAE_multimodal_endtoend  # noqa: F401,E402
from multi_modal_normative_modeling_b200.zoo import (  # noqa: F401,E402
    DMVAE, ProductOfExperts2, VariationalDecoder, VariationalEncoder, WeightedDMVAE, mmVAEPlus)
