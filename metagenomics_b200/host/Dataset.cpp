// Dataset.cpp -- see Dataset.h. The heavy lifting (parse, filter, canonical strand, parallel sort,
// dedupe) is ogb_dataset_* in libogb; this file keeps the reference's object model around it.
#include "Dataset.h"


Dataset::Dataset(void)
	: numberOfReads(0), numberOfUniqueReads(0), minimumOverlapLength(0), reads(new vector<Read *>), store(NULL), mateContext(NULL),
	  shortestReadLength(0XFFFFFFFFFFFFFFFF), longestReadLength(0)
{
}

Dataset::Dataset(vector<string> pairedEndFileNames, vector<string> singleEndFileNames, UINT64 minOverlap)
	: numberOfReads(0), numberOfUniqueReads(0), minimumOverlapLength(minOverlap), reads(new vector<Read *>), store(NULL), mateContext(NULL),
	  shortestReadLength(0XFFFFFFFFFFFFFFFF), longestReadLength(0)
{
	pairedEndDatasetFileNames = pairedEndFileNames;
	singleEndDatasetFileNames = singleEndFileNames;
	ogbCheck(ogb_dataset_create(&store), "Dataset");
	for (size_t i = 0; i < pairedEndDatasetFileNames.size(); i++)		// Dataset.cpp:51-54
		ogbCheck(ogb_dataset_add_file(store, pairedEndDatasetFileNames[i].c_str()), "Dataset::readDataset");
	for (size_t i = 0; i < singleEndDatasetFileNames.size(); i++)		// Dataset.cpp:56-59
		ogbCheck(ogb_dataset_add_file(store, singleEndDatasetFileNames[i].c_str()), "Dataset::readDataset");
	adopt(minOverlap);
}

Dataset::Dataset(const char *bases, const uint64_t *offsets, UINT64 n, UINT64 minOverlap)
	: numberOfReads(0), numberOfUniqueReads(0), minimumOverlapLength(minOverlap), reads(new vector<Read *>), store(NULL), mateContext(NULL),
	  shortestReadLength(0XFFFFFFFFFFFFFFFF), longestReadLength(0)
{
	ogbCheck(ogb_dataset_create(&store), "Dataset");
	ogbCheck(ogb_dataset_add_reads(store, bases, offsets, n), "Dataset");
	adopt(minOverlap);
}

void Dataset::adopt(UINT64 minOverlap)
{
	ogbCheck(ogb_dataset_finalize(store, (uint32_t)minOverlap), "Dataset::sortReads");	// :62-63
	numberOfReads = ogb_dataset_n_reads(store);
	numberOfUniqueReads = ogb_dataset_n_unique(store);
	if (numberOfReads) {
		shortestReadLength = ogb_dataset_shortest(store);
		longestReadLength = ogb_dataset_longest(store);
	}
	const uint16_t *len = ogb_dataset_lengths(store);
	const uint32_t *freq = ogb_dataset_frequencies(store);
	reads->reserve(numberOfUniqueReads);
	for (UINT64 i = 0; i < numberOfUniqueReads; i++) {
		Read *r = new Read;
		r->owner = this;
		r->length = len[i];
		r->setFrequency(freq[i]);
		r->setReadNumber(i + 1);											// :335-341
		reads->push_back(r);
	}
}

Dataset::~Dataset(void)
{
	for (size_t i = 0; i < reads->size(); i++) delete reads->at(i);
	delete reads;
	ogb_dataset_destroy(store);
}

UINT64 Dataset::getNumberOfReads(void) { return numberOfReads; }
UINT64 Dataset::getNumberOfUniqueReads(void) { return numberOfUniqueReads; }

string Dataset::readString(UINT64 ID, int strand) const
{
	// sized from the read's own length (a fixed 64 KB buffer per call cost ~100 GB of memset at config 2's mate-pair pass)
	uint32_t len = 0;
	const uint64_t n = ogb_dataset_n_unique(store);
	const uint32_t cap = ID >= 1 && ID <= n ? ogb_dataset_lengths(store)[ID - 1] : 0;
	string s(cap ? cap : 1, '\0');
	ogbCheck(ogb_dataset_get_read(store, ID, strand, &s[0], (uint32_t)s.size(), &len), "Dataset::getReadFromID");
	s.resize(len);
	return s;
}

Read *Dataset::getReadFromID(UINT64 ID)
{
	if (ID < 1 || ID > numberOfUniqueReads) {								// Dataset.cpp:484-490
		stringstream ss;
		ss << "ID " << ID << " out of bound.";
		throw OgbFailure(OGB_E_ARG, ss.str());
	}
	return reads->at(ID - 1);
}

Read *Dataset::getReadFromString(const string &read)
{
	uint64_t id = 0;
	ogbCheck(ogb_dataset_find_read(store, read.c_str(), (uint32_t)read.size(), &id), "Dataset::getReadFromString");
	if (id == 0) throw OgbFailure(OGB_E_ARG, "String not found in Dataset: " + read);	// Dataset.cpp:454
	return reads->at(id - 1);
}

void Dataset::readMatePairsFromFile(void)
{
	for (UINT64 i = 0; i < pairedEndDatasetFileNames.size(); i++)			// Dataset.cpp:99-102
		storeMatePairInformation(pairedEndDatasetFileNames.at(i), minimumOverlapLength, i);
}

// Dataset.cpp:208-310: re-reads a paired file two records at a time; both mates must pass the
// filter; contained reads are replaced by their super reads (:280-284); orientation bit = 1 when
// the sequence as sequenced is a substring of the stored forward strand (:291-292).
bool Dataset::storeMatePairInformation(string fileName, UINT64 minOverlap, UINT64 datasetNumber)
{
	ifstream f(fileName.c_str());
	if (!f) throw OgbFailure(OGB_E_IO, "Unable to open file: " + fileName);
	string line;
	vector<string> seqs;
	if (!getline(f, line)) return true;
	bool fasta = line[0] == '>';
	if (fasta) {
		string cur;
		while (getline(f, line)) {
			if (!line.empty() && line[0] == '>') { seqs.push_back(cur); cur.clear(); continue; }
			while (!line.empty() && (line[line.size() - 1] == '\r' || line[line.size() - 1] == '\n')) line.erase(line.size() - 1);
			cur += line;
		}
		seqs.push_back(cur);
	} else {
		for (;;) {
			string s, plus, qual;
			if (!getline(f, s)) break;
			getline(f, plus); getline(f, qual);
			while (!s.empty() && (s[s.size() - 1] == '\r')) s.erase(s.size() - 1);
			seqs.push_back(s);
			if (!getline(f, line)) break;
		}
	}
	ogb_dataset *probe = store;
	// Device path: every sequence of the file in one ogb_mate_lookup call (filter, getReadFromString, super-read redirection and
	// orientation bit on the GPU); the host only appends to the lists, in file order like the reference.
	vector<uint32_t> ids;
	vector<uint8_t> orients;
	if (mateContext != NULL && !seqs.empty()) {
		vector<uint64_t> offs(seqs.size() + 1, 0);
		for (size_t i = 0; i < seqs.size(); i++) offs[i + 1] = offs[i] + seqs[i].size();
		string flat;
		flat.reserve(offs.back());
		for (size_t i = 0; i < seqs.size(); i++) flat += seqs[i];
		ids.assign(seqs.size(), 0); orients.assign(seqs.size(), 0);
		ogbCheck(ogb_mate_lookup(mateContext, flat.data(), offs.data(), seqs.size(), (uint32_t)minOverlap, ids.data(), orients.data()), "storeMatePairInformation");
	}
	for (size_t p = 0; p + 1 < seqs.size(); p += 2) {
		if (!ids.empty() && seqs[p].size() <= 960 && seqs[p + 1].size() <= 960) {
			if (ids[p] == 0 || ids[p + 1] == 0) continue;							// one mate failed the quality filter
			Read *r1 = reads->at(ids[p] - 1), *r2 = reads->at(ids[p + 1] - 1);
			const UINT16 o1 = orients[p], o2 = orients[p + 1];
			r1->addMatePair(r2, o1 * 2 + o2, datasetNumber);
			r2->addMatePair(r1, o1 + o2 * 2, datasetNumber);
			continue;
		}
		string a = seqs[p], b = seqs[p + 1];
		for (size_t k = 0; k < a.size(); k++) a[k] = (char)toupper(a[k]);
		for (size_t k = 0; k < b.size(); k++) b[k] = (char)toupper(b[k]);
		uint64_t ia = 0, ib = 0;
		if (a.size() <= minOverlap || b.size() <= minOverlap) continue;
		ogbCheck(ogb_dataset_find_read(probe, a.c_str(), (uint32_t)a.size(), &ia), "storeMatePairInformation");
		ogbCheck(ogb_dataset_find_read(probe, b.c_str(), (uint32_t)b.size(), &ib), "storeMatePairInformation");
		if (ia == 0 || ib == 0) continue;									// one mate failed the quality filter
		Read *r1 = reads->at(ia - 1), *r2 = reads->at(ib - 1);
		if (r1->superReadID != 0) r1 = getReadFromID(r1->superReadID);
		if (r2->superReadID != 0) r2 = getReadFromID(r2->superReadID);
		UINT16 o1 = r1->getStringForward().find(a) != string::npos ? 1 : 0;
		UINT16 o2 = r2->getStringForward().find(b) != string::npos ? 1 : 0;
		r1->addMatePair(r2, o1 * 2 + o2, datasetNumber);
		r2->addMatePair(r1, o1 + o2 * 2, datasetNumber);
	}
	return true;
}

// Dataset.cpp:71-90
void Dataset::saveReads(string fileName)
{
	ofstream out(fileName.c_str());
	if (!out) throw OgbFailure(OGB_E_IO, "Unable to open file: " + fileName);
	for (UINT64 i = 1; i <= numberOfUniqueReads; i++) {
		Read *r = getReadFromID(i);
		out << setw(10) << i << (r->superReadID != 0 ? " Contained in " : " Noncontained ") << setw(10) << r->superReadID << " " << r->getStringForward() << endl;
	}
}

bool Dataset::printDataset(void)
{
	cout << "Number of reads: " << getNumberOfReads() << endl << "Number of unique reads: " << getNumberOfUniqueReads() << endl;
	for (UINT64 i = 0; i < reads->size() && i < 20; i++)
		cout << setw(10) << reads->at(i)->getReadNumber() << " " << reads->at(i)->getStringForward() << setw(10) << reads->at(i)->getFrequency() << endl;
	return true;
}
