"""Import-path shim: ``genrec.models.{hstu,sasrec,rqvae}`` resolve to the B200-native modules in ``genrec_b200`` so the
reference's gin files (``import genrec.models.hstu``, config/hstu/amazon.gin:5-6) and trainers run unchanged."""
