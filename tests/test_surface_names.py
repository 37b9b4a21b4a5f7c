"""every public name of the reference's class surface for the hot path exists on the
mirrors, with the reference's argument names (CPU: no device call is made)."""
import inspect

LHMM_NAMES = ("states t transmat B_p pi profunction ksai_acc gamma_acc fix_code add_acc cal_observation_pro "
              "save_parameter save_acc init_parameter init_acc clear_result_buffer clear_data change_t add_T change_pi "
              "change_A update_acc update_param baulm_welch viterbi init_data add_data data datasize dimension classes "
              "markdata").split()
GMM_NAMES = ("mean covariance alpha dimension mixture data acc alpha_acc mean_acc covariance_acc bias gmm_id add_data "
             "clear_data save_parameter save_acc init_parameter init_acc expectation q_function maximization update_acc "
             "update_param em theta gmm point").split()
CI_NAMES = "cal_distance cal_variance kmeans".split()
AM_NAMES = ("loaded_units statenum load_unit init_unit init_parameter delete_buffer_file delete_trainInfo init_audio "
            "process_data multi_process_data training multi_training embedded_training multi_embedded_training_1 "
            "multi_embedded_training_2 discriminate embedded viterbi VirtualState").split()


def test_every_public_name_exists():
    from poccala_b200.AcousticModel import AcousticModel
    from poccala_b200.Clustering import Clustering
    from poccala_b200.LHMM import LHMM

    for cls, names in ((LHMM, LHMM_NAMES), (Clustering.GMM, GMM_NAMES), (Clustering.ClusterInitialization, CI_NAMES),
                       (AcousticModel, AM_NAMES)):
        missing = [n for n in names if not hasattr(cls, n)]
        assert not missing, (cls.__name__, missing)


def test_signatures_match_the_reference():
    from poccala_b200.AcousticModel import AcousticModel
    from poccala_b200.Clustering import Clustering
    from poccala_b200.LHMM import LHMM

    def args(f):
        return [p for p in inspect.signature(f).parameters if p != "self"]

    assert args(LHMM.__init__) == ["states", "statesnum", "log", "t", "transmat", "profunc", "probmat", "pi", "hmm_list",
                                   "fix_code"]
    assert args(LHMM.viterbi) == ["log", "states", "transmat", "prob", "pi", "convert", "end_state_back", "show_mark_state"]
    assert args(LHMM.init_data) == ["data", "datapath", "shuffle", "continuous", "matrix", "hasmark"]
    assert args(Clustering.GMM.__init__) == ["log", "dimension", "mix_level", "data", "alpha", "mean", "variance",
                                             "covariance", "differentiation", "gmm_id"]
    assert args(Clustering.GMM.em) == ["show_q", "smem", "c_covariance"]
    assert args(Clustering.ClusterInitialization.__init__) == ["data", "k", "dimension", "log"]
    assert args(AcousticModel.__init__)[:11] == ["log", "unit_type", "mode", "processes", "job_id", "console", "state_num",
                                                 "mix_level", "dct_num", "delta_1", "delta_2"]
    assert args(AcousticModel.multi_process_data) == ["label", "data", "init", "args"]
    assert args(AcousticModel.init_audio) == ["audiopath", "labelpath"]


def test_init_data_reads_the_reference_text_format(tmp_path):
    from poccala_b200.LHMM import LHMM
    import numpy as np

    p = tmp_path / "d.csv"
    p.write_text("toy data\n4 2 2 red white\n1.0,2.0,red\n3.0,4.0,white\n5.0,6.0,red\n7.5,8.5,white\n")
    h = LHMM({0: "a"}, 5, None, probmat=[np.zeros((1, 1))])
    h.init_data(datapath=str(p))
    assert h.datasize == 4 and h.dimension == 2 and h.classes == 2
    assert sorted(h.markdata) == ["red", "white"] and len(h.markdata["red"]) == 2
    assert np.allclose(h.data[3], [7.5, 8.5])


def test_init_audio_lists_a_corpus_directory(tmp_path):
    from poccala_b200.AcousticModel import AcousticModel

    (tmp_path / "wav").mkdir()
    for n in ("A1_0", "A1_1"):
        (tmp_path / "wav" / (n + ".wav")).write_bytes(b"")
    it = AcousticModel.init_audio(str(tmp_path / "wav"), "/lab")
    assert next(it) == 2
    pairs = sorted(it)
    assert pairs[0] == (str(tmp_path / "wav") + "/A1_0.wav\n", "/lab/A1_0.wav.trn\n")
