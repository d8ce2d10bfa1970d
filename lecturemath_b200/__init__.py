"""lecturemath_b200 -- B200-native (sm_100a) implementation of LectureMath's per-frame content-extraction hot path.

Host-side mirrors of the reference classes over the C ABI of libaccessmath_b200.so (include/accessmath_b200.h):
    fcn_lecturenet.FCN_LectureNet, fcn_binarizer_worker.FCN_LectureNet_Binarizer, labeler.Labeler,
    connected_component.ConnectedComponent, cc_stability_estimator.CCStabilityEstimator (+ the stage-03 methods of cc_grouping),
    helper.Helper, configuration.Configuration, and the fused pipeline (pipeline.ContentExtractor / StreamingExtractor).
Nothing is imported eagerly: building plans and loading the library need no device, running does (there is no CPU fallback)."""
__version__ = "0.1.0"
